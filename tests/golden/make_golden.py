"""Generate golden vectors by running the UNMODIFIED reference (torch CPU fp64) in the build container.

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

Reads /root/reference (read-only; exists only in the build container, never on the GPU box) and writes
tests/golden/*.npz.  `cyipopt` is not installed, so an empty stub module is injected before importing
`src.mpc` (only `get_optimal_trajectory` touches it).  All random inputs are seeded here; the reference
itself has no seeds.
"""
import os
import sys
import types

import numpy as np
import torch

sys.dont_write_bytecode = True
REF = "/root/reference"
sys.path.insert(0, REF)
sys.modules.setdefault("cyipopt", types.ModuleType("cyipopt"))

from src.gpr import GaussianProcessRegression  # noqa: E402
from src.dynamics import Dynamics  # noqa: E402
from src.mpc import RiskSensitiveMPC  # noqa: E402
from src.tools.uncertainty_prop import (mean_prop_torch, variance_prop_torch, covariance_prop_torch,  # noqa: E402
                                        mean_prop, variance_prop, covariance_prop)

OUT = os.path.dirname(os.path.abspath(__file__))
T = lambda a: torch.tensor(np.asarray(a, dtype=np.float64))  # noqa: E731


def synth(n, E, m, seed):
    """SURVEY 8d generator: contracting tanh dynamics on the unit box."""
    rng = np.random.default_rng(seed)
    D = E + m
    S = rng.uniform(-1, 1, (n, E))
    A = rng.uniform(-1, 1, (n, m))
    Wm = rng.normal(0, 0.3, (D, E))
    nxt = 0.9 * S + 0.2 * np.tanh(np.concatenate([S, A], 1) @ Wm)
    return S, A, nxt, rng


def gpr_cases():
    out = {}
    rng = np.random.default_rng(11)
    for name, n, D, lam, sf, sn in [("a", 40, 3, [0.7, 1.3, 2.0], 1.2, 0.3), ("b", 97, 5, [2.0] * 5, 1.0, 0.1),
                                    ("c", 1, 2, [1.0, 1.0], 1.0, 1.0)]:
        X = rng.normal(size=(n, D))
        y = np.sin(X.sum(1)) + 0.1 * rng.normal(size=n)
        g = GaussianProcessRegression(D)
        g.set_lambdas(lam); g.set_sigma_f(sf); g.set_sigma_n(sn)
        g.append_train_data(X if n > 1 else X[0], y if n > 1 else float(y[0]))
        Xp = rng.normal(size=(7, D))
        mean, cov = g.predict_latent_vars(Xp, covar=True, targets=False)
        mean_t, cov_t = g.predict_latent_vars(Xp, covar=True, targets=True)
        Kpt = g.compute_pred_train_covariance(Xp).detach().numpy()
        k1 = g.compute_pred_train_covariance(Xp[0]).detach().numpy()
        out.update({f"gpr_{name}_X": X, f"gpr_{name}_y": y, f"gpr_{name}_lam": np.array(lam),
                    f"gpr_{name}_sf": sf, f"gpr_{name}_sn": sn, f"gpr_{name}_Kf": g.Kf.detach().numpy(),
                    f"gpr_{name}_lam_eff": g.get_lambdas(), f"gpr_{name}_sf_eff": g.get_sigma_f(),
                    f"gpr_{name}_sn_eff": g.get_sigma_n(),
                    f"gpr_{name}_Ky": g.Ky.detach().numpy(), f"gpr_{name}_Kyinv": g.Ky_inv.detach().numpy(),
                    f"gpr_{name}_Xp": Xp, f"gpr_{name}_mean": mean, f"gpr_{name}_cov": cov,
                    f"gpr_{name}_cov_targets": cov_t, f"gpr_{name}_Kpt": Kpt, f"gpr_{name}_k1": k1})
    return out


def mm_cases():
    """mean/variance/covariance propagation with FULL input covariance, ARD lambdas, sigma_f != 1."""
    out = {}
    rng = np.random.default_rng(5)
    for name, n, D, sf, sn in [("a", 60, 2, 1.0, 0.5), ("b", 150, 3, 1.3, 0.2), ("c", 120, 5, 0.8, 0.1)]:
        X = rng.normal(size=(n, D)) * 1.2
        y = (X ** 2).sum(1) * 0.3 + 0.1 * rng.normal(size=n)
        lam1 = rng.uniform(0.6, 2.5, D)
        lam2 = rng.uniform(0.6, 2.5, D)
        u = rng.normal(size=D) * 0.5
        Am = rng.normal(size=(D, D)) * 0.3
        S = Am @ Am.T + 0.05 * np.eye(D)
        res = {}
        for tag, lam in (("1", lam1), ("2", lam2)):
            g = GaussianProcessRegression(D)
            g.set_lambdas(lam); g.set_sigma_f(sf); g.set_sigma_n(sn)
            g.append_train_data(X, y)
            Kinv = g.Ky_inv.detach()
            mu, dd = mean_prop_torch(Kinv, T(lam), T(u), T(S), T(X), T(y), sf)
            var = variance_prop_torch(Kinv, T(lam), T(u), T(S), T(X), mu, dd["beta"], sf)
            Sd = np.diag(np.diag(S))
            mud, ddd = mean_prop_torch(Kinv, T(lam), T(u), T(Sd), T(X), T(y), sf)
            vard = variance_prop_torch(Kinv, T(lam), T(u), T(Sd), T(X), mud, ddd["beta"], sf)
            assert np.array_equal(g.get_lambdas(), lam)   # fp64 ndarray in -> no fp32 rounding
            res[tag] = (Kinv.numpy(), mu.item(), var.item(), dd["beta"].numpy(), dd["l"].numpy(), mud.item(), vard.item())
        cov_t = covariance_prop_torch(T(lam1), T(lam2), T(u), T(S), T(X), res["1"][1], res["2"][1],
                                      T(res["1"][3]), T(res["2"][3]), sf, sf).item()
        out.update({f"mm_{name}_X": X, f"mm_{name}_y": y, f"mm_{name}_lam1": lam1, f"mm_{name}_lam2": lam2,
                    f"mm_{name}_u": u, f"mm_{name}_S": S, f"mm_{name}_sf": sf, f"mm_{name}_sn": sn,
                    f"mm_{name}_Kinv1": res["1"][0], f"mm_{name}_Kinv2": res["2"][0],
                    f"mm_{name}_mean1": res["1"][1], f"mm_{name}_var1": res["1"][2],
                    f"mm_{name}_beta1": res["1"][3], f"mm_{name}_l1": res["1"][4],
                    f"mm_{name}_mean2": res["2"][1], f"mm_{name}_var2": res["2"][2],
                    f"mm_{name}_beta2": res["2"][3],
                    f"mm_{name}_mean1_diag": res["1"][5], f"mm_{name}_var1_diag": res["1"][6],
                    f"mm_{name}_cov12_torch": cov_t})
        if name == "a":
            # sigma_f = 1 NumPy twins (they take K, not K^-1, and assume sigma_f = 1)
            K1 = np.linalg.inv(res["1"][0]); K2 = np.linalg.inv(res["2"][0])
            out[f"mm_{name}_cov12_numpy"] = covariance_prop(K1, K2, np.diag(lam1), np.diag(lam2), u, S, X, y)
            out[f"mm_{name}_mean1_numpy"] = mean_prop(K1, np.diag(lam1), u, S, X, y)[0]
            out[f"mm_{name}_var1_numpy"] = variance_prop(K1, np.diag(lam1), u, S, X, y)
    return out


def make_mpc(S, A, nxt, lam, sf, sn, gamma, H, Q, R, Rd=None):
    E = S.shape[1]; m = A.shape[1]
    mpc = RiskSensitiveMPC(gamma, H, E, m, Q, R, Rd)
    for a in range(E):
        mpc.dynamics.gpr_err[a].set_sigma_n(sn[a])
        mpc.dynamics.gpr_err[a].set_lambdas(list(lam[a]))
        mpc.dynamics.gpr_err[a].set_sigma_f(sf[a])
    mpc.dynamics.append_train_data(S, A, nxt)
    return mpc


def rollout_cases():
    out = {}
    # (name, n, E, m, H, gamma, per-output ARD?, R_delta?)
    specs = [("r1", 64, 2, 1, 3, -1.0, False, False), ("r2", 200, 4, 1, 5, -1.0, False, False),
             ("r3", 150, 3, 2, 4, 0.7, True, True), ("r4", 512, 4, 1, 4, -1.0, False, False),
             ("r5", 96, 1, 1, 6, -0.5, True, False)]
    for name, n, E, m, H, gamma, ard, use_rd in specs:
        S, A, nxt, rng = synth(n, E, m, seed=hash(name) % 1000 if False else sum(map(ord, name)))
        D = E + m
        if ard:
            lam = rng.uniform(1.0, 3.0, (E, D)); sf = rng.uniform(0.8, 1.3, E); sn = rng.uniform(0.08, 0.2, E)
            Qm = rng.normal(size=(E, E)) * 0.3; Q = Qm @ Qm.T + 1.5 * np.eye(E)
            Rm = rng.normal(size=(m, m)) * 0.1; R = Rm @ Rm.T + 0.05 * np.eye(m)
        else:
            lam = np.full((E, D), 2.0); sf = np.ones(E); sn = np.full(E, 0.1)
            Q = 2.0 * np.eye(E); R = 0.01 * np.eye(m)
        Rd = (0.3 * np.eye(m) + 0.05) if use_rd else None
        mpc = make_mpc(S, A, nxt, lam, sf, sn, gamma, H, Q, R, Rd)
        xref = rng.uniform(-0.2, 0.2, E); uref = rng.uniform(-0.1, 0.1, m)
        mpc.set_xref(xref); mpc.set_uref(uref)
        last = rng.uniform(-0.3, 0.3, H * m)
        mpc.last_traj = last
        x0 = rng.uniform(-0.5, 0.5, E)
        U = rng.uniform(-0.3, 0.3, (H, m))
        mpc.curr_state = T(x0)
        c = mpc.objective(U.reshape(-1).copy())
        g = np.array(mpc.gradient(U.reshape(-1).copy()))
        means, covs = mpc.dynamics.forward_propagate_torch(H, T(x0), T(U))
        gp = mpc.dynamics.gpr_err
        lam = np.stack([gp[a].get_lambdas() for a in range(E)])          # effective values
        sf = np.array([gp[a].get_sigma_f() for a in range(E)])
        sn = np.array([gp[a].get_sigma_n() for a in range(E)])
        out.update({f"{name}_S": S, f"{name}_A": A, f"{name}_next": nxt, f"{name}_lam": lam, f"{name}_sf": sf,
                    f"{name}_sn": sn, f"{name}_gamma": gamma, f"{name}_Q": Q, f"{name}_R": R,
                    f"{name}_Rd": Rd if Rd is not None else np.zeros((0, 0)), f"{name}_last": last,
                    f"{name}_xref": xref, f"{name}_uref": uref, f"{name}_x0": x0, f"{name}_U": U,
                    f"{name}_cost": c, f"{name}_grad": g,
                    f"{name}_means": torch.stack(means).detach().numpy(),
                    f"{name}_covs": torch.stack(covs).detach().numpy()})
        print(name, "cost", c)
    return out


def shipped_case():
    """BASELINE config 1 inputs: src/experiments/pretrain_uncertainty.py:87-113 (shipped .npy data)."""
    d = os.path.join(REF, "src", "experiments", "data")
    S = np.load(os.path.join(d, "states.npy")); A = np.load(os.path.join(d, "actions.npy"))
    nxt = np.load(os.path.join(d, "next_states.npy"))
    H = 6
    mpc = make_mpc(S, A, nxt, np.full((2, 4), 0.5), np.ones(2), np.full(2, 1e-5), -1, H, 2 * np.eye(2),
                   np.zeros((2, 2)))
    mpc.set_xref(np.zeros(2)); mpc.set_uref(np.zeros(2))
    x0 = np.array([4.0, -4.0])
    mpc.curr_state = T(x0)
    gp = mpc.dynamics.gpr_err
    out = {"ship_S": S, "ship_A": A, "ship_next": nxt, "ship_x0": x0,
           "ship_lam": np.stack([gp[a].get_lambdas() for a in range(2)]),
           "ship_sf": np.array([gp[a].get_sigma_f() for a in range(2)]),
           "ship_sn": np.array([gp[a].get_sigma_n() for a in range(2)])}
    rng = np.random.default_rng(3)
    Us = [np.zeros((H, 2)), np.full((H, 2), -0.5), rng.uniform(-0.2, 0.2, (H, 2)),
          np.stack([np.full(H, -0.15), np.full(H, 0.15)], 1)]
    for i, U in enumerate(Us):
        c = mpc.objective(U.reshape(-1).copy())
        g = np.array(mpc.gradient(U.reshape(-1).copy()))
        means, covs = mpc.dynamics.forward_propagate_torch(H, T(x0), T(U))
        out.update({f"ship_U{i}": U, f"ship_cost{i}": c, f"ship_grad{i}": g,
                    f"ship_means{i}": torch.stack(means).detach().numpy(),
                    f"ship_covs{i}": torch.stack(covs).detach().numpy()})
        print("shipped", i, c)
    return out


def hyper_cases():
    """Marginal likelihood, its autograd gradient and three Adam steps (src/gpr.py:240-251,334-370)."""
    import io, contextlib
    out = {}
    rng = np.random.default_rng(21)
    n, D = 60, 3
    X = rng.normal(size=(n, D)); y = np.sin(X @ np.array([1.0, 0.5, -0.7])) + 0.1 * rng.normal(size=n)
    g = GaussianProcessRegression(D)
    g.set_lambdas(np.array([0.8, 1.5, 2.2])); g.set_sigma_f(np.float64(1.3)); g.set_sigma_n(np.float64(0.25))
    # the optimiser of the reference keeps the tensors created in __init__: rebuild it on the new ones
    g.optimizer = torch.optim.Adam(params=[g.log_lambdas, g.log_sigma_n, g.log_sigma_f], lr=0.1, betas=(0.9, 0.999),
                                   maximize=True)
    g.append_train_data(X, y)
    ml = g.compute_marginal_likelihood()
    ml.backward()
    out.update({"hy_X": X, "hy_y": y, "hy_lam0": g.get_lambdas(), "hy_sf0": g.get_sigma_f(), "hy_sn0": g.get_sigma_n(),
                "hy_ml": ml.item(), "hy_dlam": g.log_lambdas.grad.numpy().copy(),
                "hy_dsf": g.log_sigma_f.grad.item(), "hy_dsn": g.log_sigma_n.grad.item()})
    g.optimizer.zero_grad()
    g.build_Ky_inv_mat()
    with contextlib.redirect_stdout(io.StringIO()):
        g.update_hyperparams(num_iters=3)
    out.update({"hy_lam3": g.get_lambdas(), "hy_sf3": g.get_sigma_f(), "hy_sn3": g.get_sigma_n(),
                "hy_ml3": g.compute_marginal_likelihood().item()})
    return out


def cost_kats():
    """Deterministic known-answer tests of the reference, src/test/test_mpc.py:15-57,245-274."""
    out = {}
    mpc = RiskSensitiveMPC(1, 1, 2, 2, np.array([[2, 0], [0, 2]]), np.array([[1, 1], [1, 1]]))
    x = np.array([[1., 1.], [3., 3.]]); u = np.array([[2., 2.]])
    sig = np.array([[[1., 2.], [3., 4.]], [[5., 6.], [7., 8.]]])
    out["kat_cost"] = mpc.cost(x, u, sig, np.array([0.5, 0.5]), np.array([0.6, 0.6]))
    out["kat_cost_torch"] = mpc.cost_torch(T(x), T(u), T(sig), T([0.5, 0.5]), T([0.6, 0.6])).item()
    H = 5
    mpc = RiskSensitiveMPC(-1, H, 1, 1, 2 * np.identity(1), np.array([[0]]), np.array([[0]]))
    xt = T([5, 4, 3, 2, 1, 0]).reshape(H + 1, 1)
    st = T([1 / 6, 1 / 7, 1 / 8, 1 / 9, 1 / 10, 1 / 11]).reshape(H + 1, 1, 1)
    out["kat_state_cost"] = mpc.cost_torch(xt, torch.zeros((H, 1)).type(torch.float64), st,
                                           torch.zeros(1).type(torch.float64),
                                           torch.zeros(1).type(torch.float64)).item()
    return out


def fullcov_cases():
    """Full-covariance rollouts assembled from the reference's NumPy moment-matching functions
    (`src/tools/uncertainty_prop.py:6-44,91-136,187-236`) exactly as the block the reference left commented out would
    (`src/dynamics.py:104-121`), and the reference's NumPy cost on the resulting full Sigma (`src/mpc.py:118-154`).
    `covariance_prop` takes ONE target vector for both GPs, so both outputs are trained on the same targets and differ
    in their hyper-parameters: fc1 shares the length-scales (different noise), fc2 has distinct ARD length-scales."""
    out = {}
    for name, n, lam, sn, seed in [("fc1", 60, [[1.5, 1.5, 1.5]] * 2, [0.1, 0.3], 41),
                                   ("fc2", 50, [[1.2, 2.2, 0.8], [2.0, 0.9, 1.6]], [0.2, 0.2], 42)]:
        E, m, H = 2, 1, 3
        rng = np.random.default_rng(seed)
        X = rng.uniform(-1, 1, (n, E + m))
        y = 0.8 * X[:, 0] - 0.3 * X[:, 1] + 0.2 * np.tanh(X @ rng.normal(0, 0.5, E + m)) + 0.05 * rng.normal(size=n)
        lam = np.array(lam, dtype=np.float64)
        gps = []
        for a in range(E):
            g = GaussianProcessRegression(E + m)
            g.set_lambdas(np.asarray(lam[a], dtype=np.float64)); g.set_sigma_f(np.float64(1.0)); g.set_sigma_n(np.float64(sn[a]))
            g.append_train_data(X, y)
            gps.append(g)
        Ks = [g.Ky.detach().numpy() for g in gps]
        Lams = [np.diag(lam[a]) for a in range(E)]
        x0 = rng.uniform(-0.4, 0.4, E); U = rng.uniform(-0.3, 0.3, (H, m))
        means = np.zeros((H + 1, E)); covs = np.zeros((H + 1, E, E))
        means[0] = x0; covs[0] = 1e-3 * np.identity(E)
        act = (1e-3 * torch.eye(m)).type(torch.float64).numpy()      # fp32 eye promoted, src/dynamics.py:162
        for t in range(1, H + 1):
            mean_in = np.concatenate((means[t - 1], U[t - 1]))
            covar = np.zeros((E + m, E + m)); covar[:E, :E] = covs[t - 1]; covar[E:, E:] = act
            for a in range(E):
                means[t, a] = mean_prop(Ks[a], Lams[a], mean_in, covar, X, y)[0]
                covs[t, a, a] = variance_prop(Ks[a], Lams[a], mean_in, covar, X, y)
            for i in range(1, E):
                for j in range(i):
                    covs[t, i, j] = covs[t, j, i] = covariance_prop(Ks[i], Ks[j], Lams[i], Lams[j], mean_in, covar, X, y)
        Q = np.array([[2.0, 0.3], [0.3, 1.5]]); R = 0.05 * np.eye(m)
        xref = np.array([0.1, -0.05]); uref = np.array([0.02])
        costs = {}
        for gamma in (-1.0, 0.7):
            mpc = RiskSensitiveMPC(gamma, H, E, m, Q, R)
            costs[gamma] = mpc.cost(means, U, covs, xref, uref)
        out.update({f"{name}_X": X, f"{name}_y": y, f"{name}_lam": lam, f"{name}_sn": np.array([g.get_sigma_n() for g in gps]),
                    f"{name}_x0": x0, f"{name}_U": U, f"{name}_means": means, f"{name}_covs": covs, f"{name}_Q": Q,
                    f"{name}_R": R, f"{name}_xref": xref, f"{name}_uref": uref,
                    f"{name}_cost_gm1": costs[-1.0], f"{name}_cost_gp07": costs[0.7]})
        print(name, "cov[H]", covs[H].tolist(), "cost", costs)
    return out


def fnom_case():
    """predict_latent_vars with a nominal model, `src/gpr.py:305-309` (residual form)."""
    rng = np.random.default_rng(13)
    n, D = 45, 3
    X = rng.normal(size=(n, D)); y = X[:, 0] * 0.6 + np.sin(X.sum(1)) * 0.3 + 0.05 * rng.normal(size=n)
    f_nom = lambda Z: 0.5 * Z[:, :1] + 0.1          # noqa: E731  (n, 1), like y_train
    g = GaussianProcessRegression(D, nominal_model=f_nom)
    g.set_lambdas(np.array([0.9, 1.4, 2.0])); g.set_sigma_f(np.float64(1.1)); g.set_sigma_n(np.float64(0.15))
    g.append_train_data(X, y)
    Xp = rng.normal(size=(6, D))
    mean, cov = g.predict_latent_vars(Xp, covar=True, targets=True)
    return {"fn_X": X, "fn_y": y, "fn_lam": g.get_lambdas(), "fn_sf": g.get_sigma_f(), "fn_sn": g.get_sigma_n(),
            "fn_Xp": Xp, "fn_mean": mean, "fn_cov_targets": cov}


if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count())
    np.savez_compressed(os.path.join(OUT, "fullcov.npz"), **fullcov_cases())
    np.savez_compressed(os.path.join(OUT, "gpr_fnom.npz"), **fnom_case())
    if "--new-only" in sys.argv:
        sys.exit(0)
    np.savez_compressed(os.path.join(OUT, "gpr.npz"), **gpr_cases())
    np.savez_compressed(os.path.join(OUT, "moment_matching.npz"), **mm_cases())
    np.savez_compressed(os.path.join(OUT, "rollout.npz"), **rollout_cases())
    np.savez_compressed(os.path.join(OUT, "shipped.npz"), **shipped_case())
    np.savez_compressed(os.path.join(OUT, "cost_kat.npz"), **cost_kats())
    np.savez_compressed(os.path.join(OUT, "hyper.npz"), **hyper_cases())
    print("golden vectors written to", OUT)
