"""Generates tests/golden/hybrid_rows.npz: a small seeded window with inverse-depth features and what the oracle
restatement (oracle/hybrid.py) computes for it -- H1 per-observation Jacobians, H2 stacked rows + gate, H4 anchor change.
The reference has no test data for these functions; the fixture pins the oracle (tests/test_oracle_hybrid_cpu.py checks
that it still reproduces it) and gives the GPU tests a target that does not import the oracle at run time.

    python tests/golden/make_hybrid_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import hybrid as hy          # noqa: E402
from oracle import mathutils as mu       # noqa: E402


def build():
    rng = np.random.default_rng(2024)
    N, F = 8, 6
    R_b2c, t_c_b = mu.so3_exp(rng.normal(0, 0.8, 3)), rng.normal(0, 0.1, 3)
    clone_R = np.array([mu.so3_exp(rng.normal(0, 0.08, 3)) for _ in range(N)])
    clone_p = np.array([np.array([0.25 * i, 0.02 * i, 0.0]) + rng.normal(0, 0.03, 3) for i in range(N)])
    anchor = rng.integers(0, N - 1, F).astype(np.int32)
    rho = 1.0 / rng.uniform(3.0, 20.0, F)
    f_an = np.stack([rng.uniform(-0.4, 0.4, F), rng.uniform(-0.3, 0.3, F)], axis=1)
    pos = np.array([hy.feature_position_from_anchor(clone_R[a], clone_p[a], R_b2c, t_c_b, [fx, fy, 1.0], r)
                    for a, r, (fx, fy) in zip(anchor, rho, f_an)])
    feat_off, obs_clone, obs_z = [0], [], []
    for f in range(F):
        cl = np.sort(rng.choice(N, int(rng.integers(2, N + 1)), replace=False))
        for c in cl:
            p_ck = R_b2c @ clone_R[c].T @ (pos[f] - (clone_p[c] + clone_R[c] @ t_c_b))
            obs_clone.append(int(c))
            obs_z.append(p_ck[:2] / p_ck[2] + rng.normal(0, 0.004, 2))
        feat_off.append(len(obs_clone))
    feat_off, obs_clone, obs_z = np.array(feat_off, np.int32), np.array(obs_clone, np.int32), np.array(obs_z)
    H_f, H_a, H_x, H_e, r = [], [], [], [], []
    for f in range(F):
        a = int(anchor[f])
        for o in range(feat_off[f], feat_off[f + 1]):
            c = int(obs_clone[o])
            out = hy.measurement_jacobian_ekf_1didp(clone_R[c], clone_p[c], clone_R[a], clone_p[a], R_b2c, t_c_b,
                                                    [f_an[f, 0], f_an[f, 1], 1.0], rho[f], pos[f], obs_z[o], same_state=(c == a))
            for lst, v in zip((H_f, H_a, H_x, H_e, r), out):
                lst.append(np.asarray(v))
    D = 22 + 6 * N + F
    A = rng.normal(0, 0.03, (D, D))
    P = A @ A.T * 0.1 + 1e-5 * np.eye(D)
    P[15:22, :] = 0.0
    P[:, 15:22] = 0.0
    k = N - 1
    z_cur = np.array([(lambda p_ck: p_ck[:2] / p_ck[2])(R_b2c @ clone_R[k].T @ (pos[f] - (clone_p[k] + clone_R[k] @ t_c_b)))
                      for f in range(F)]) + rng.normal(0, 0.01, (F, 2))
    sigma2 = 6.4e-5
    chi2 = mu.chi2_table(0.95)[2]
    rows_H, rows_r, gamma = [], [], []
    for f in range(F):
        H, rr = hy.feature_jacobian_ekf(clone_R, clone_p, R_b2c, t_c_b, k, int(anchor[f]), f, F,
                                        [f_an[f, 0], f_an[f, 1], 1.0], rho[f], pos[f], z_cur[f])
        g, _ = hy.gate_ekf_row(H, rr, P, sigma2, chi2)
        rows_H.append(H)
        rows_r.append(rr)
        gamma.append(g)
    fidx, old, new = 2, int(anchor[2]), N - 1
    p_c = R_b2c @ clone_R[new].T @ (pos[fidx] - (clone_p[new] + clone_R[new] @ t_c_b))
    rho_new = 1.0 / p_c[2]
    P_re, J_re = hy.update_feature_cov_1didp(P, N, fidx, old, new, clone_R, clone_p, R_b2c, t_c_b, pos[fidx], rho_new)
    return dict(clone_R=clone_R.reshape(N, 9), clone_p=clone_p, R_b2c=R_b2c, t_c_b=t_c_b, anchor=anchor, inv_depth=rho,
                f_an=f_an, positions=pos, feat_off=feat_off, obs_clone=obs_clone, obs_z=obs_z,
                H_f=np.array(H_f).reshape(-1, 2), H_a=np.array(H_a), H_x=np.array(H_x), H_e=np.array(H_e),
                r=np.array(r), P=P, z_cur=z_cur, sigma2=sigma2, rows_H=np.vstack(rows_H), rows_r=np.concatenate(rows_r),
                gamma=np.array(gamma), chi2=chi2, re_fidx=fidx, re_old=old, re_new=new, re_rho_new=rho_new, P_re=P_re,
                J_re=J_re[0])


if __name__ == "__main__":
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "hybrid_rows.npz")
    np.savez_compressed(out, **build())
    print("wrote", out, os.path.getsize(out), "bytes")
