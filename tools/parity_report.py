"""Measured deviations of the CUDA path from the reference's golden numbers (tests/golden/rollout.npz: objective,
autograd gradient, rollout means / variances of the UNMODIFIED reference), as a table for DESIGN.md.
    python tools/parity_report.py
"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import gpmpc_b200 as gp

g = np.load(os.path.join(ROOT, "tests", "golden", "rollout.npz"))
T = lambda a: torch.tensor(np.asarray(a, dtype=np.float64), device="cuda:0")     # noqa: E731
print("| case | n | E | m | H | cost rel. err | gradient err / max | means err / max | variances err / max(|ref|, 1e-3) |")
print("|---|---|---|---|---|---|---|---|---|")
for name in ("r1", "r2", "r3", "r4", "r5"):
    S, A = g[f"{name}_S"], g[f"{name}_A"]
    E, m = S.shape[1], A.shape[1]
    U = g[f"{name}_U"]; H = U.shape[0]
    Rd = g[f"{name}_Rd"]; Rd = None if Rd.size == 0 else Rd
    mpc = gp.RiskSensitiveMPC(float(g[f"{name}_gamma"]), H, E, m, g[f"{name}_Q"], g[f"{name}_R"], Rd)
    for a in range(E):
        mpc.dynamics.gpr_err[a].set_lambdas(np.asarray(g[f"{name}_lam"][a], dtype=np.float64))
        mpc.dynamics.gpr_err[a].set_sigma_f(np.float64(g[f"{name}_sf"][a]))
        mpc.dynamics.gpr_err[a].set_sigma_n(np.float64(g[f"{name}_sn"][a]))
    mpc.dynamics.append_train_data(S, A, g[f"{name}_next"])
    mpc.set_xref(g[f"{name}_xref"]); mpc.set_uref(g[f"{name}_uref"])
    mpc.last_traj = g[f"{name}_last"]
    mpc.curr_state = T(g[f"{name}_x0"])
    c = mpc.objective(U.reshape(-1).copy()); grad = np.asarray(mpc.gradient(U.reshape(-1).copy()))
    means, covs = mpc.dynamics.forward_propagate(H, g[f"{name}_x0"], U)
    cref = float(g[f"{name}_cost"]); gref = g[f"{name}_grad"]; mref = g[f"{name}_means"]; vref = np.diagonal(g[f"{name}_covs"], axis1=1, axis2=2)
    v = np.diagonal(covs, axis1=1, axis2=2)
    print(f"| {name} | {S.shape[0]} | {E} | {m} | {H} | {abs(c - cref) / abs(cref):.1e} | {np.max(np.abs(grad - gref)) / np.max(np.abs(gref)):.1e} | "
          f"{np.max(np.abs(means - mref)) / np.max(np.abs(mref)):.1e} | {np.max(np.abs(v - vref)) / max(np.max(np.abs(vref)), 1e-3):.1e} |")
