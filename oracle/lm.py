"""CPU oracle -- the reference's Levenberg-Marquardt (TEST INFRASTRUCTURE, not product code).

Restatement of the vendored MINPACK port include/orcvio/utils/EigenLevenbergMarquardt/ (LevenbergMarquardt.h:300-395
minimize / minimizeInit / lmder1, LMonestep.h:22-206, LMpar.h:20-158 lmpar2, LMqrsolv.h:22-103), generalised the way
the reference generalised it: the unknown lives on a manifold (`plus(x, dx)`, `scaled_norm(diag, x)`).
Eigen's ColPivHouseholderQR is restated by LAPACK's pivoted QR (scipy.linalg.qr(pivoting=True): same pivot rule,
largest remaining column norm).

parity: pinned by the reference's own known answers src/tests/test_levenberg_marquardt.cpp:64-140 (lmder1: info 1,
nfev 6, njev 5, |f| = 0.09063596, x; the quadratic: status CosinusTooSmall, nfev 2, njev 2) -- tests/test_oracle_cpu.py.
"""
import math

import numpy as np
import scipy.linalg

# LevenbergMarquardtSpace::Status
RUNNING, IMPROPER, REL_RED_TOO_SMALL, REL_ERR_TOO_SMALL, REL_BOTH_TOO_SMALL, COS_TOO_SMALL, TOO_MANY_FEV, \
    FTOL_TOO_SMALL, XTOL_TOO_SMALL, GTOL_TOO_SMALL, USER_ASKED = -1, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9
EPS = np.finfo(float).eps
DWARF = np.finfo(float).tiny


def _qrsolv(s, perm, diag, qtb):
    """lmqrsolv, LMqrsolv.h:22-103.  s: n x n (upper triangle = R, modified below the diagonal); returns (x, sdiag)."""
    n = s.shape[0]
    x = np.diag(s).copy()
    wa = qtb.copy()
    for i in range(n):
        for j in range(i):
            s[i, j] = s[j, i]
    sdiag = np.zeros(n)
    for j in range(n):
        l = perm[j]
        if diag[l] == 0.0:
            break
        sdiag[j:] = 0.0
        sdiag[j] = diag[l]
        qtbpj = 0.0
        for k in range(j, n):
            # JacobiRotation::makeGivens(p = -s(k,k), q = sdiag[k]) (Eigen/src/Jacobi/Jacobi.h, real case)
            p, q = -s[k, k], sdiag[k]
            if q == 0.0:
                c, sn = (-1.0 if p < 0 else 1.0), 0.0
            elif p == 0.0:
                c, sn = 0.0, (1.0 if q < 0 else -1.0)
            elif abs(p) > abs(q):
                t = q / p
                u = math.sqrt(1.0 + t * t)
                if p < 0:
                    u = -u
                c = 1.0 / u
                sn = -t * c
            else:
                t = p / q
                u = math.sqrt(1.0 + t * t)
                if q < 0:
                    u = -u
                sn = -1.0 / u
                c = -t * sn
            s[k, k] = c * s[k, k] + sn * sdiag[k]
            temp = c * wa[k] + sn * qtbpj
            qtbpj = -sn * wa[k] + c * qtbpj
            wa[k] = temp
            for i in range(k + 1, n):
                temp = c * s[i, k] + sn * sdiag[i]
                sdiag[i] = -sn * s[i, k] + c * sdiag[i]
                s[i, k] = temp
    nsing = 0
    while nsing < n and sdiag[nsing] != 0.0:
        nsing += 1
    wa[nsing:] = 0.0
    if nsing:
        wa[:nsing] = scipy.linalg.solve_triangular(s[:nsing, :nsing].T, wa[:nsing], lower=False)
    sdiag = np.diag(s).copy()
    for i in range(n):
        s[i, i] = x[i]
    out = np.zeros(n)
    out[perm] = wa                      # x = P * wa
    return out, sdiag


def _lmpar2(R, perm, rank, diag, qtb, delta, par):
    """lmpar2, LMpar.h:20-158.  Returns (par, x)."""
    n = R.shape[0]
    s = R.copy()
    wa1 = qtb.copy()
    wa1[rank:] = 0.0
    if rank:
        wa1[:rank] = scipy.linalg.solve_triangular(s[:rank, :rank], qtb[:rank], lower=False)
    x = np.zeros(n)
    x[perm] = wa1
    wa2 = diag * x
    dxnorm = np.linalg.norm(wa2)
    fp = dxnorm - delta
    if fp <= 0.1 * delta:
        return 0.0, x
    parl = 0.0
    if rank == n:
        wa1 = (diag * wa2 / dxnorm)[perm]
        wa1 = scipy.linalg.solve_triangular(s.T, wa1, lower=True)
        temp = np.linalg.norm(wa1)
        parl = fp / delta / temp / temp
    wa1 = np.array([s[:j + 1, j] @ qtb[:j + 1] / diag[perm[j]] for j in range(n)])
    gnorm = np.linalg.norm(wa1)
    paru = gnorm / delta
    if paru == 0.0:
        paru = DWARF / min(delta, 0.1)
    par = max(par, parl)
    par = min(par, paru)
    if par == 0.0:
        par = gnorm / dxnorm
    it = 0
    while True:
        it += 1
        if par == 0.0:
            par = max(DWARF, 0.001 * paru)
        wa1 = math.sqrt(par) * diag
        x, sdiag = _qrsolv(s, perm, wa1, qtb)
        wa2 = diag * x
        dxnorm = np.linalg.norm(wa2)
        temp = fp
        fp = dxnorm - delta
        if abs(fp) <= 0.1 * delta or (parl == 0.0 and fp <= temp and temp < 0.0) or it == 10:
            break
        wa1 = (diag * (wa2 / dxnorm))[perm]
        for j in range(n):
            wa1[j] /= sdiag[j]
            temp = wa1[j]
            for i in range(j + 1, n):
                wa1[i] -= s[i, j] * temp
        temp = np.linalg.norm(wa1)
        parc = fp / delta / temp / temp
        if fp > 0.0:
            parl = max(parl, par)
        if fp < 0.0:
            paru = min(paru, par)
        par = max(parl, par + parc)
    return par, x


def minimize(fun, jac, x0, plus=None, scaled_norm=None, ftol=None, xtol=None, gtol=0.0, factor=100.0, maxfev=400):
    """LevenbergMarquardt::minimize.  fun(x) -> residuals (m), jac(x) -> m x n.  Returns dict(x, status, nfev, njev,
    fnorm, iterations)."""
    ftol = math.sqrt(EPS) if ftol is None else ftol
    xtol = math.sqrt(EPS) if xtol is None else xtol
    plus = plus or (lambda x, d: x + d)
    scaled_norm = scaled_norm or (lambda diag, x: float(np.linalg.norm(diag * x)))
    x = x0
    fvec = np.asarray(fun(x), dtype=float)
    m = fvec.shape[0]
    nfev, njev, it, par = 1, 0, 1, 0.0
    fnorm = float(np.linalg.norm(fvec))
    diag = delta = xnorm = None
    n = None
    status = RUNNING
    while status == RUNNING:
        J = np.asarray(jac(x), dtype=float)
        njev += 1
        n = J.shape[1]
        if n <= 0 or m < n:
            return dict(x=x, status=IMPROPER, nfev=nfev, njev=njev, fnorm=fnorm, iterations=it)
        wa2 = np.linalg.norm(J, axis=0)
        Q, R, perm = scipy.linalg.qr(J, mode="economic", pivoting=True)
        R = R[:n, :n]
        dR = np.abs(np.diag(R))
        rank = int(np.sum(dR > dR.max() * EPS * min(m, n))) if dR.size and dR.max() > 0 else 0
        if it == 1:
            diag = np.where(wa2 == 0.0, 1.0, wa2)
            xnorm = scaled_norm(diag, x)
            delta = factor * xnorm
            if delta == 0.0:
                delta = factor
        qtf = (Q.T @ fvec)[:n]
        gnorm = 0.0
        if fnorm != 0.0:
            for j in range(n):
                if wa2[perm[j]] != 0.0:
                    gnorm = max(gnorm, abs(R[:j + 1, j] @ (qtf[:j + 1] / fnorm) / wa2[perm[j]]))
        if gnorm <= gtol:
            status = COS_TOO_SMALL
            break
        diag = np.maximum(diag, wa2)
        while True:
            par, wa1 = _lmpar2(R, perm, rank, diag, qtf, delta, par)
            wa1 = -wa1
            x_try = plus(x, wa1)
            pnorm = float(np.linalg.norm(diag * wa1))
            if it == 1:
                delta = min(delta, pnorm)
            f_try = np.asarray(fun(x_try), dtype=float)
            nfev += 1
            fnorm1 = float(np.linalg.norm(f_try))
            actred = -1.0
            if 0.1 * fnorm1 < fnorm:
                actred = 1.0 - (fnorm1 / fnorm) ** 2
            wa3 = R @ wa1[perm]
            temp1 = (float(np.linalg.norm(wa3)) / fnorm) ** 2
            temp2 = (math.sqrt(par) * pnorm / fnorm) ** 2
            prered = temp1 + temp2 / 0.5
            dirder = -(temp1 + temp2)
            ratio = actred / prered if prered != 0.0 else 0.0
            if ratio <= 0.25:
                temp = 0.5 if actred >= 0.0 else 0.5 * dirder / (dirder + 0.5 * actred)
                if 0.1 * fnorm1 >= fnorm or temp < 0.1:
                    temp = 0.1
                delta = temp * min(delta, pnorm / 0.1)
                par /= temp
            elif not (par != 0.0 and ratio < 0.75):
                delta = pnorm / 0.5
                par = 0.5 * par
            if ratio >= 1e-4:
                x, fvec = x_try, f_try
                xnorm = scaled_norm(diag, x)
                fnorm = fnorm1
                it += 1
            small = abs(actred) <= ftol and prered <= ftol and 0.5 * ratio <= 1.0
            if small and delta <= xtol * xnorm:
                status = REL_BOTH_TOO_SMALL
            elif small:
                status = REL_RED_TOO_SMALL
            elif delta <= xtol * xnorm:
                status = REL_ERR_TOO_SMALL
            elif nfev >= maxfev:
                status = TOO_MANY_FEV
            elif abs(actred) <= EPS and prered <= EPS and 0.5 * ratio <= 1.0:
                status = FTOL_TOO_SMALL
            elif delta <= EPS * xnorm:
                status = XTOL_TOO_SMALL
            elif gnorm <= EPS:
                status = GTOL_TOO_SMALL
            if status != RUNNING or ratio >= 1e-4:
                break
    return dict(x=x, status=status, nfev=nfev, njev=njev, fnorm=fnorm, iterations=it)


def lmder1(fun, jac, x0, tol=None):
    """LevenbergMarquardt::lmder1 (LevenbergMarquardt.h:377-395)."""
    tol = math.sqrt(EPS) if tol is None else tol
    n = len(x0)
    return minimize(fun, jac, x0, ftol=tol, xtol=tol, gtol=0.0, factor=100.0, maxfev=100 * (n + 1))


# the functor of src/tests/test_levenberg_marquardt.cpp:27-61 (MINPACK's lmder1 example)
_Y = np.array([1.4e-1, 1.8e-1, 2.2e-1, 2.5e-1, 2.9e-1, 3.2e-1, 3.5e-1, 3.9e-1, 3.7e-1, 5.8e-1, 7.3e-1, 9.6e-1, 1.34, 2.1, 4.39])


def kat_fun(x):
    i = np.arange(15)
    t1, t2 = i + 1.0, 16.0 - i - 1.0
    t3 = np.where(i >= 8, t2, t1)
    return _Y - (x[0] + t1 / (x[1] * t2 + x[2] * t3))


def kat_jac(x):
    i = np.arange(15)
    t1, t2 = i + 1.0, 16.0 - i - 1.0
    t3 = np.where(i >= 8, t2, t1)
    t4 = (x[1] * t2 + x[2] * t3) ** 2
    return np.column_stack([-np.ones(15), t1 * t2 / t4, t1 * t3 / t4])
