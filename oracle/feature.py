"""CPU oracle -- per-feature triangulation (TEST INFRASTRUCTURE, not product code).

Restates Feature::triangulate_position and its helpers,
include/orcvio/feat/feature.hpp:271-351 (cost / jacobian / generateInitialGuess),
:353-396 (checkMotion), :583-719 (LM loop and validity tests).

Written with plain Python floats in one fixed operation order (IEEE double, no FMA)
because the accept/reject decisions of the LM loop (`new_cost < total_cost`,
`delta_norm > 5e-7`) are rounding sensitive: the C restatement (oracle/cpu_ref.cpp)
and the CUDA kernel (orcvio_b200/csrc/tri_kernel.cu, compiled --fmad=false) follow the
same order term for term, so all three agree bit for bit.

parity: UNPINNED by the reference (no reference test covers feature.hpp).
Pinned instead by oracle == C restatement == CUDA and by reprojection properties.
"""
import math
from types import SimpleNamespace


def default_opt_config():
    """Feature::OptimizationConfig defaults, feature.hpp:51-59."""
    return SimpleNamespace(translation_threshold=0.2, huber_epsilon=0.01,
                           estimation_precision=5e-7, initial_damping=1e-3,
                           outer_loop_max_iteration=10, inner_loop_max_iteration=10,
                           cost_threshold=4.7673e-04, init_final_dist_threshold=5.0)


def _matT_mat(A, B):
    """A^T * B for row-major 3x3 lists; entry = (a0*b0 + a1*b1) + a2*b2."""
    out = [0.0] * 9
    for i in range(3):
        for j in range(3):
            out[3 * i + j] = (A[0 + i] * B[0 + j] + A[3 + i] * B[3 + j]) + A[6 + i] * B[6 + j]
    return out


def _matT_vec(A, v):
    return [(A[0 + i] * v[0] + A[3 + i] * v[1]) + A[6 + i] * v[2] for i in range(3)]


def _mat_vec(A, v):
    return [(A[3 * i] * v[0] + A[3 * i + 1] * v[1]) + A[3 * i + 2] * v[2] for i in range(3)]


def ldlt3_solve(M, b):
    """Pivoted LDL^T solve of a symmetric 3x3 system, structured like
    Eigen::LDLT<Matrix3d>::compute/solve (third party, Eigen 3.3; call site
    feature.hpp:649): largest-|diagonal| pivoting, in-place lower storage,
    forward / diagonal / backward substitution."""
    n = 3
    a = [[M[3 * i + j] for j in range(n)] for i in range(n)]
    tr = [0] * n
    for k in range(n):
        big = k
        bigv = abs(a[k][k])
        for i in range(k + 1, n):
            if abs(a[i][i]) > bigv:
                bigv = abs(a[i][i])
                big = i
        tr[k] = big
        if big != k:
            for j in range(k):
                a[k][j], a[big][j] = a[big][j], a[k][j]
            for i in range(big + 1, n):
                a[i][k], a[i][big] = a[i][big], a[i][k]
            a[k][k], a[big][big] = a[big][big], a[k][k]
            for i in range(k + 1, big):
                a[i][k], a[big][i] = a[big][i], a[i][k]
        if k > 0:
            temp = [a[j][j] * a[k][j] for j in range(k)]
            s = 0.0
            for j in range(k):
                s = s + a[k][j] * temp[j]
            a[k][k] = a[k][k] - s
            for i in range(k + 1, n):
                s = 0.0
                for j in range(k):
                    s = s + a[i][j] * temp[j]
                a[i][k] = a[i][k] - s
        akk = a[k][k]
        if abs(akk) > 0.0:
            for i in range(k + 1, n):
                a[i][k] = a[i][k] / akk
    x = list(b)
    for k in range(n):
        if tr[k] != k:
            x[k], x[tr[k]] = x[tr[k]], x[k]
    # L y = x (unit lower, column oriented)
    for i in range(n):
        for r in range(i + 1, n):
            x[r] = x[r] - x[i] * a[r][i]
    # D
    tol = 2.2250738585072014e-308
    for i in range(n):
        if abs(a[i][i]) > tol:
            x[i] = x[i] / a[i][i]
        else:
            x[i] = 0.0
    # L^T (unit upper, dot product then subtract)
    for i in range(n - 2, -1, -1):
        s = 0.0
        for j in range(i + 1, n):
            s = s + a[j][i] * x[j]
        x[i] = x[i] - s
    for k in range(n - 1, -1, -1):
        if tr[k] != k:
            x[k], x[tr[k]] = x[tr[k]], x[k]
    return x


def _h(R, t, x):
    a, b, rho = x
    return [((R[0] * a + R[1] * b) + R[2] * 1.0) + rho * t[0],
            ((R[3] * a + R[4] * b) + R[5] * 1.0) + rho * t[1],
            ((R[6] * a + R[7] * b) + R[8] * 1.0) + rho * t[2]]


def cost(R, t, x, z):
    """Feature::cost, feature.hpp:271-291."""
    h = _h(R, t, x)
    d0 = h[0] / h[2] - z[0]
    d1 = h[1] / h[2] - z[1]
    return d0 * d0 + d1 * d1


def jacobian(R, t, x, z, huber_epsilon):
    """Feature::jacobian, feature.hpp:293-329 -> (J[2][3], r[2], w)."""
    h = _h(R, t, x)
    W = [[R[0], R[1], t[0]], [R[3], R[4], t[1]], [R[6], R[7], t[2]]]
    ih3 = 1 / h[2]
    c0 = h[0] / (h[2] * h[2])
    c1 = h[1] / (h[2] * h[2])
    J = [[ih3 * W[0][j] - c0 * W[2][j] for j in range(3)],
         [ih3 * W[1][j] - c1 * W[2][j] for j in range(3)]]
    r = [h[0] / h[2] - z[0], h[1] / h[2] - z[1]]
    e = math.sqrt(r[0] * r[0] + r[1] * r[1])
    if e <= huber_epsilon:
        w = 1.0
    else:
        w = math.sqrt(2.0 * huber_epsilon / e)
    return J, r, w


def generate_initial_guess(R, t, z1, z2):
    """Feature::generateInitialGuess, feature.hpp:331-351."""
    m = _mat_vec(R, [z1[0], z1[1], 1.0])
    A0 = m[0] - z2[0] * m[2]
    A1 = m[1] - z2[1] * m[2]
    b0 = z2[0] * t[2] - t[0]
    b1 = z2[1] * t[2] - t[1]
    inv = 1.0 / (A0 * A0 + A1 * A1)
    depth = (inv * A0) * b0 + (inv * A1) * b1
    return [z1[0] * depth, z1[1] * depth, depth]


def check_motion(R_first, t_first, t_last, z_first, translation_threshold):
    """Feature::checkMotion, feature.hpp:353-396 (poses are cam->world)."""
    d = [z_first[0], z_first[1], 1.0]
    n = math.sqrt((d[0] * d[0] + d[1] * d[1]) + d[2] * d[2])
    d = [d[0] / n, d[1] / n, d[2] / n]
    d = _mat_vec(R_first, d)
    tr = [t_last[i] - t_first[i] for i in range(3)]
    par = (tr[0] * d[0] + tr[1] * d[1]) + tr[2] * d[2]
    o = [tr[i] - par * d[i] for i in range(3)]
    return math.sqrt((o[0] * o[0] + o[1] * o[1]) + o[2] * o[2]) > translation_threshold


def triangulate(cam_R, cam_t, meas, is_initialized, position, cfg):
    """Feature::triangulate_position, feature.hpp:583-719.

    cam_R[i]: row-major 3x3 list (camera i -> world), cam_t[i]: 3-list,
    meas[i]: (u, v).  Returns SimpleNamespace(valid, solution, final_position,
    position (world), R_last, t_last, n_outer, n_inner_total, total_cost)."""
    m = len(cam_R)
    Rl = list(cam_R[m - 1])
    tl = list(cam_t[m - 1])
    rel_R, rel_t = [], []
    for i in range(m):
        Ri = cam_R[i]
        tinv = _matT_vec(Ri, cam_t[i])
        tinv = [-tinv[0], -tinv[1], -tinv[2]]
        rel_R.append(_matT_mat(Ri, Rl))
        rt = _matT_vec(Ri, tl)
        rel_t.append([rt[0] + tinv[0], rt[1] + tinv[1], rt[2] + tinv[2]])

    if not is_initialized:
        init = generate_initial_guess(rel_R[0], rel_t[0], meas[m - 1], meas[0])
    else:
        tinv = _matT_vec(Rl, tl)
        rp = _matT_vec(Rl, position)
        init = [rp[0] + (-tinv[0]), rp[1] + (-tinv[1]), rp[2] + (-tinv[2])]
    sol = [init[0] / init[2], init[1] / init[2], 1.0 / init[2]]

    lam = cfg.initial_damping
    inner = 0
    outer = 0
    reduced = False
    delta_norm = 0.0
    n_inner_total = 0
    total_cost = 0.0
    for i in range(m):
        total_cost = total_cost + cost(rel_R[i], rel_t[i], sol, meas[i])

    while True:
        A = [0.0] * 9
        b = [0.0] * 3
        for i in range(m):
            J, r, w = jacobian(rel_R[i], rel_t[i], sol, meas[i], cfg.huber_epsilon)
            if w == 1:
                for a_ in range(3):
                    for c_ in range(3):
                        A[3 * a_ + c_] = A[3 * a_ + c_] + (J[0][a_] * J[0][c_] + J[1][a_] * J[1][c_])
                    b[a_] = b[a_] + (J[0][a_] * r[0] + J[1][a_] * r[1])
            else:
                w2 = w * w
                for a_ in range(3):
                    for c_ in range(3):
                        A[3 * a_ + c_] = A[3 * a_ + c_] + ((w2 * J[0][a_]) * J[0][c_] + (w2 * J[1][a_]) * J[1][c_])
                    b[a_] = b[a_] + ((w2 * J[0][a_]) * r[0] + (w2 * J[1][a_]) * r[1])
        while True:
            M = list(A)
            M[0] = A[0] + lam
            M[4] = A[4] + lam
            M[8] = A[8] + lam
            delta = ldlt3_solve(M, b)
            new_sol = [sol[0] - delta[0], sol[1] - delta[1], sol[2] - delta[2]]
            delta_norm = math.sqrt((delta[0] * delta[0] + delta[1] * delta[1]) + delta[2] * delta[2])
            new_cost = 0.0
            for i in range(m):
                new_cost = new_cost + cost(rel_R[i], rel_t[i], new_sol, meas[i])
            n_inner_total += 1
            if new_cost < total_cost:
                reduced = True
                sol = new_sol
                total_cost = new_cost
                lam = lam / 10 if lam / 10 > 1e-10 else 1e-10
            else:
                reduced = False
                lam = lam * 10 if lam * 10 < 1e12 else 1e12
            cont = (inner < cfg.inner_loop_max_iteration) and (not reduced)
            inner += 1
            if not cont:
                break
        inner = 0
        cont = (outer < cfg.outer_loop_max_iteration) and (delta_norm > cfg.estimation_precision)
        outer += 1
        if not cont:
            break

    final = [sol[0] / sol[2], sol[1] / sol[2], 1.0 / sol[2]]
    valid = True
    for i in range(m):
        R = rel_R[i]
        pz = ((R[6] * final[0] + R[7] * final[1]) + R[8] * final[2]) + rel_t[i][2]
        if pz <= 0:
            valid = False
            break
    normalized_cost = total_cost / float(2 * m * m)
    d = [final[0] - init[0], final[1] - init[1], final[2] - init[2]]
    if math.sqrt((d[0] * d[0] + d[1] * d[1]) + d[2] * d[2]) > cfg.init_final_dist_threshold:
        valid = False
    if normalized_cost > cfg.cost_threshold:
        valid = False
    pw = _mat_vec(Rl, final)
    pw = [pw[0] + tl[0], pw[1] + tl[1], pw[2] + tl[2]]
    return SimpleNamespace(valid=valid, solution=sol, final_position=final, position=pw,
                           R_last=Rl, t_last=tl, n_outer=outer, n_inner_total=n_inner_total,
                           total_cost=total_cost, init=init)
