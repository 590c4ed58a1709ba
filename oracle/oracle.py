"""CPU oracle for the GP-MPC rollout hot path -- TEST INFRASTRUCTURE ONLY.

This module is a plain NumPy (fp64) restatement of the algorithm the reference implements in
`src/gpr.py`, `src/tools/uncertainty_prop.py`, `src/dynamics.py` and `src/mpc.py`.  It is the checker
the CUDA path is compared against.  Only `tests/`, `__graft_entry__.smoke()` and the CPU-baseline legs
of `bench.py` may import it; the product package never does.

Parity status: PINNED.  Every function here is checked in `tests/test_oracle.py` against golden vectors
produced by running the unmodified reference (torch CPU fp64) in the build container
(`tests/golden/make_golden.py` -> `tests/golden/*.npz`) and against the reference's deterministic
known-answer tests (`src/test/test_mpc.py:15-57,245-274`).

All citations are relative to /root/reference.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

# The reference builds the action block of the input covariance from an fp32 eye
# (`src/dynamics.py:162`), so the action variance is the fp32 rounding of 1e-3 promoted to fp64.
ACTION_VAR = float(np.float32(1e-3))
# The initial state covariance is an exact fp64 1e-3 (`src/dynamics.py:148`).
STATE0_VAR = 1e-3


# --------------------------------------------------------------------------------------------
# GP regression (src/gpr.py)
# --------------------------------------------------------------------------------------------
def se_gram(X1, X2, lambdas, sigma_f):
    """Squared-exponential ARD kernel matrix, `src/gpr.py:124-135,167-169,273-276`.

    lambdas are *squared* length-scales: k(x,x') = sf^2 exp(-1/2 sum_k (x_k-x'_k)^2 / lambda_k).
    """
    X1 = np.atleast_2d(np.asarray(X1, dtype=np.float64))
    X2 = np.atleast_2d(np.asarray(X2, dtype=np.float64))
    lam = np.asarray(lambdas, dtype=np.float64)
    d = X1[:, None, :] - X2[None, :, :]
    return sigma_f ** 2 * np.exp(-0.5 * np.sum(d * d / lam, axis=2))


def fit(X, y, lambdas, sigma_f, sigma_n):
    """Kf, Ky, Ky_inv, beta as in `src/gpr.py:159-171` (+ beta = Ky_inv y, `uncertainty_prop.py:327`).

    The reference forms the explicit LU inverse; so does the oracle.
    """
    X = np.asarray(X, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64).reshape(-1)
    Kf = se_gram(X, X, lambdas, sigma_f)
    Ky = Kf + sigma_n ** 2 * np.eye(X.shape[0])
    Ky_inv = np.linalg.inv(Ky)
    beta = Ky_inv @ y
    return {"Kf": Kf, "Ky": Ky, "Ky_inv": Ky_inv, "beta": beta}


def predict(X, y, lambdas, sigma_f, sigma_n, X_pred, covar=False, targets=False, Ky_inv=None):
    """Posterior mean (p,1) and covariance (p,p), `src/gpr.py:285-332` with f_nom = None."""
    X = np.asarray(X, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64).reshape(-1, 1)
    Xp = np.atleast_2d(np.asarray(X_pred, dtype=np.float64))
    if Ky_inv is None:
        Ky_inv = fit(X, y, lambdas, sigma_f, sigma_n)["Ky_inv"]
    Ks = se_gram(Xp, X, lambdas, sigma_f)
    mean = Ks @ Ky_inv @ y
    if not covar:
        return mean, None
    cov = se_gram(Xp, Xp, lambdas, sigma_f) - Ks @ Ky_inv @ Ks.T
    if targets:
        cov = cov + sigma_n ** 2 * np.eye(Xp.shape[0])
    return mean, cov


def marginal_likelihood(X, y, lambdas, sigma_f, sigma_n):
    """Log marginal likelihood, `src/gpr.py:240-247` (f_nom = None): -1/2 y^T Ky^-1 y - 1/2 log det Ky - n/2 log 2pi."""
    f = fit(X, y, lambdas, sigma_f, sigma_n)
    y = np.asarray(y, dtype=np.float64).reshape(-1)
    return float(-0.5 * y @ f["Ky_inv"] @ y - 0.5 * np.log(np.linalg.det(f["Ky"])) - 0.5 * len(y) * np.log(2 * np.pi))


# --------------------------------------------------------------------------------------------
# Moment matching (src/tools/uncertainty_prop.py) -- general (full) input covariance S
# --------------------------------------------------------------------------------------------
def mean_prop(Ky_inv, lambdas, u, S, X, y, sigma_f=1.0):
    """Exact mean of the GP output for x* ~ N(u,S): `uncertainty_prop.py:296-338`.

    Returns (mean, beta, l).
    """
    lam = np.asarray(lambdas, dtype=np.float64)
    u = np.asarray(u, dtype=np.float64)
    S = np.asarray(S, dtype=np.float64)
    beta = Ky_inv @ np.asarray(y, dtype=np.float64).reshape(-1)
    d = S.shape[0]
    B = np.linalg.inv(S + np.diag(lam))
    V = u[None, :] - X
    quad = np.sum((V @ B) * V, axis=1)
    det = np.linalg.det(np.diag(1.0 / lam) @ S + np.eye(d))
    l = det ** (-0.5) * np.exp(-0.5 * quad) * sigma_f ** 2
    return float(beta @ l), beta, l


def variance_prop(Ky_inv, lambdas, u, S, X, mean, beta, sigma_f=1.0):
    """Exact variance of the GP output for x* ~ N(u,S): `uncertainty_prop.py:341-399`.

    The reference evaluates trace((Ky_inv - beta beta^T) @ L); because L is symmetric this equals the
    element-wise sum  sum_ij (Ky_inv - beta beta^T)_ij L_ij  that is evaluated here.
    """
    lam = np.asarray(lambdas, dtype=np.float64)
    u = np.asarray(u, dtype=np.float64)
    S = np.asarray(S, dtype=np.float64)
    d = S.shape[0]
    A = np.linalg.inv(np.diag(lam) / 2.0 + S)
    det = np.linalg.det(2.0 * np.diag(1.0 / lam) @ S + np.eye(d)) ** (-0.5)
    V = u[None, :] - X                       # v_i = u - x_i
    q = V @ A @ V.T                          # q_ij = v_i^T A v_j
    qd = np.diag(q)
    a_part = np.exp(-0.125 * (qd[:, None] + 2.0 * q + qd[None, :]))
    Xs = X / np.sqrt(lam)
    dd = Xs[:, None, :] - Xs[None, :, :]
    lam_part = np.exp(-0.25 * np.sum(dd * dd, axis=2))
    L = det * a_part * lam_part * sigma_f ** 4
    W = Ky_inv - np.outer(beta, beta)
    return float(sigma_f ** 2 - np.sum(W * L) - mean ** 2)


def covariance_prop(lam1, lam2, u, S, X, mean1, mean2, beta1, beta2, sigma_f1=1.0, sigma_f2=1.0,
                    bugcompat=False):
    """Cross-covariance of two GP outputs: NumPy form `uncertainty_prop.py:187-236` (correct),
    torch form `:402-465`.  `bugcompat=True` reproduces the torch function's transposed cross term
    (`:446`), which differs from the formula unless lam1 is proportional to lam2.
    """
    lam1 = np.asarray(lam1, dtype=np.float64)
    lam2 = np.asarray(lam2, dtype=np.float64)
    u = np.asarray(u, dtype=np.float64)
    S = np.asarray(S, dtype=np.float64)
    d = S.shape[0]
    R = S @ np.diag(1.0 / lam1 + 1.0 / lam2) + np.eye(d)
    det = np.linalg.det(R) ** (-0.5)
    T = np.linalg.inv(R) @ S
    Xc = X - u[None, :]
    z1 = Xc / lam1                            # rows: Lambda1^-1 (x_i - u)
    z2 = Xc / lam2
    a1 = np.sum((z1 @ T.T) * z1, axis=1)      # z1_i^T T z1_i
    a2 = np.sum((z2 @ T.T) * z2, axis=1)
    if bugcompat:
        cross = z2 @ T @ z1.T                 # [i,j] = z2_i^T T z1_j   (reference torch, :446)
    else:
        cross = z1 @ T @ z2.T                 # [i,j] = z1_i^T T z2_j   (formula / NumPy twin)
    # z_ij^T T z_ij with z_ij = z1_i + z2_j (T symmetric)
    expo = 0.5 * (a1[:, None] + 2.0 * cross + a2[None, :])
    k1 = np.sum(Xc * Xc / lam1, axis=1)
    k2 = np.sum(Xc * Xc / lam2, axis=1)
    Qt = det * np.exp(-0.5 * (k1[:, None] + k2[None, :])) * np.exp(expo) * sigma_f1 ** 2 * sigma_f2 ** 2
    return float(beta1 @ Qt @ beta2 - mean1 * mean2)


# --------------------------------------------------------------------------------------------
# Rollout (src/dynamics.py:126-191) and cost (src/mpc.py:156-200)
# --------------------------------------------------------------------------------------------
def rollout(X, Ky_invs, Ys, lambdas, sigma_fs, x0, U):
    """Variance-only moment-matched rollout.  X:[n,D]; Ky_invs: list of E [n,n]; Ys:[n,E];
    lambdas:[E,D]; sigma_fs:[E]; x0:[E]; U:[H,m].  Returns means [H+1,E], variances [H+1,E]."""
    X = np.asarray(X, dtype=np.float64)
    E = len(Ky_invs)
    U = np.asarray(U, dtype=np.float64)
    H, m = U.shape
    means = np.zeros((H + 1, E))
    vars_ = np.zeros((H + 1, E))
    means[0] = x0
    vars_[0] = STATE0_VAR
    for t in range(1, H + 1):
        u = np.concatenate([means[t - 1], U[t - 1]])
        S = np.diag(np.concatenate([vars_[t - 1], np.full(m, ACTION_VAR)]))
        for a in range(E):
            mu, beta, _ = mean_prop(Ky_invs[a], lambdas[a], u, S, X, Ys[:, a], sigma_fs[a])
            means[t, a] = mu
            vars_[t, a] = variance_prop(Ky_invs[a], lambdas[a], u, S, X, mu, beta, sigma_fs[a])
    return means, vars_


def moment_match_full(X, Ky_invs, betas, lambdas, sigma_fs, u, S, use_c=False):
    """mean [E] and the full E x E covariance of the E GP outputs for x* ~ N(u, S), S full:
    variances from `variance_prop` (`uncertainty_prop.py:341-399`), cross-covariances from the formula-correct
    NumPy form (`uncertainty_prop.py:187-236`).  `use_c` evaluates the same sums with the C restatement."""
    E = len(Ky_invs)
    if use_c:
        X = _c(X); n, D = X.shape
        Kinv = _c(np.stack(Ky_invs)); bt = _c(np.stack(betas)); lam = _c(lambdas); sf = _c(sigma_fs)
        uu = _c(u); SS = _c(S)
        mean = np.zeros(E); cov = np.zeros((E, E))
        _lib().oracle_moment_match_full(n, D, E, _p(X), _p(Kinv), _p(bt), _p(lam), _p(sf), _p(uu), _p(SS), _p(mean), _p(cov))
        return mean, cov
    mean = np.zeros(E); cov = np.zeros((E, E))
    for a in range(E):
        V = np.asarray(u)[None, :] - X
        Bm = np.linalg.inv(np.asarray(S) + np.diag(lambdas[a]))
        det = np.linalg.det(np.diag(1.0 / np.asarray(lambdas[a])) @ S + np.eye(len(u)))
        l = det ** (-0.5) * np.exp(-0.5 * np.sum((V @ Bm) * V, axis=1)) * sigma_fs[a] ** 2
        mean[a] = betas[a] @ l
    for a in range(E):
        cov[a, a] = variance_prop(Ky_invs[a], lambdas[a], u, S, X, mean[a], betas[a], sigma_fs[a])
        for b in range(a + 1, E):
            cov[a, b] = cov[b, a] = covariance_prop(lambdas[a], lambdas[b], u, S, X, mean[a], mean[b], betas[a],
                                                    betas[b], sigma_fs[a], sigma_fs[b])
    return mean, cov


def rollout_full(X, Ky_invs, betas, lambdas, sigma_fs, x0, U, use_c=False):
    """Full-covariance moment-matched rollout: like `rollout` (`src/dynamics.py:126-191`) but Sigma_t keeps the
    cross-covariances between the outputs -- the wiring the reference leaves as a TODO (`src/dynamics.py:104-121,184`).
    Input covariance of step t: blockdiag(Sigma_{t-1}, fp32(1e-3) I).  Returns means [H+1,E], covs [H+1,E,E]."""
    X = np.asarray(X, dtype=np.float64)
    E = len(Ky_invs)
    U = np.asarray(U, dtype=np.float64)
    H, m = U.shape
    means = np.zeros((H + 1, E)); covs = np.zeros((H + 1, E, E))
    means[0] = x0
    covs[0] = STATE0_VAR * np.eye(E)
    for t in range(1, H + 1):
        u = np.concatenate([means[t - 1], U[t - 1]])
        S = np.zeros((E + m, E + m))
        S[:E, :E] = covs[t - 1]
        S[E:, E:] = ACTION_VAR * np.eye(m)
        means[t], covs[t] = moment_match_full(X, Ky_invs, betas, lambdas, sigma_fs, u, S, use_c=use_c)
    return means, covs


def rollout_full_cost(X, Ky_invs, betas, lambdas, sigma_fs, x0, U, gamma, Q, R, R_delta=None, last_u=None,
                      x_ref=None, u_ref=None, use_c=False):
    """Cost of a control sequence under the full-covariance rollout (`src/mpc.py:156-200` accepts a full Sigma)."""
    means, covs = rollout_full(X, Ky_invs, betas, lambdas, sigma_fs, x0, U, use_c=use_c)
    E = means.shape[1]; m = np.asarray(U).shape[1]
    xr = np.zeros(E) if x_ref is None else np.asarray(x_ref, dtype=np.float64)
    ur = np.zeros(m) if u_ref is None else np.asarray(u_ref, dtype=np.float64)
    return cost(means, U, covs, xr, ur, gamma, Q, R, R_delta, last_u), means, covs


def cost(means, U, covs, x_ref, u_ref, gamma, Q, R, R_delta=None, last_u=None):
    """Risk-sensitive cost, `src/mpc.py:156-200`.  covs: [H+1,E,E] (full) or [H+1,E] (diagonal)."""
    means = np.asarray(means, dtype=np.float64)
    U = np.asarray(U, dtype=np.float64)
    covs = np.asarray(covs, dtype=np.float64)
    Q = np.asarray(Q, dtype=np.float64)
    R = np.asarray(R, dtype=np.float64)
    E = means.shape[1]
    H = U.shape[0]
    if covs.ndim == 2:
        covs = np.stack([np.diag(c) for c in covs])
    Qi = np.linalg.inv(Q)
    c = 0.0
    with np.errstate(invalid="ignore", divide="ignore"):
        for i in range(H + 1):
            c = c + (1.0 / gamma) * np.log(np.linalg.det(np.eye(E) + gamma * Q @ covs[i]))
            e = means[i] - x_ref
            c = c + e @ np.linalg.inv(Qi + gamma * covs[i]) @ e
    for j in range(H):
        du = U[j] - u_ref
        c = c + du @ R @ du
    if R_delta is not None:
        Rd = np.asarray(R_delta, dtype=np.float64)
        lu = np.asarray(last_u, dtype=np.float64).reshape(1, -1)
        dU = np.diff(np.concatenate([lu, U], axis=0), axis=0)
        for j in range(H):
            c = c + dU[j] @ Rd @ dU[j]
    return float(c)


# --------------------------------------------------------------------------------------------
# C restatement (oracle/gpmpc_oracle.c): O(n^2) pair sums + closed-form adjoint, OpenMP threaded.
# --------------------------------------------------------------------------------------------
_LIB = None


def build_c(force=False):
    so = os.path.join(_HERE, "libgpmpc_oracle.so")
    src = os.path.join(_HERE, "gpmpc_oracle.c")
    if force or (not os.path.exists(so)) or os.path.getmtime(so) < os.path.getmtime(src):
        gcc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"   # /opt/gcc lacks libgomp
        base = [gcc, "-O2", "-fPIC", "-shared", "-o", so, src, "-lm"]
        if subprocess.call(base[:2] + ["-fopenmp"] + base[2:]) != 0:
            subprocess.check_call(base)
    return so


def _lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "libgpmpc_oracle.so")
        if not os.path.exists(so):
            build_c()
        _LIB = ctypes.CDLL(so)
        dp = ctypes.POINTER(ctypes.c_double)
        _LIB.oracle_moment_match_diag.argtypes = [ctypes.c_int, ctypes.c_int, dp, dp, dp, dp,
                                                  ctypes.c_double, dp, dp, dp]
        _LIB.oracle_moment_match_diag.restype = None
        _LIB.oracle_rollout_cost_grad.argtypes = [ctypes.c_int] * 5 + [dp] * 7 + [ctypes.c_double] + \
                                                 [dp] * 10
        _LIB.oracle_rollout_cost_grad.restype = None
        _LIB.oracle_moment_match_full.argtypes = [ctypes.c_int] * 3 + [dp] * 9
        _LIB.oracle_moment_match_full.restype = None
        _LIB.oracle_set_threads.argtypes = [ctypes.c_int]
        _LIB.oracle_set_threads.restype = ctypes.c_int
    return _LIB


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def _c(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


def c_set_threads(k):
    return _lib().oracle_set_threads(int(k))


def c_moment_match_diag(X, Ky_inv, beta, lambdas, sigma_f, u, s):
    """(mean, var, partials[4D]) for diagonal input variance s; see oracle_moment_match_diag."""
    X = _c(X); Ky_inv = _c(Ky_inv); beta = _c(beta); lambdas = _c(lambdas); u = _c(u); s = _c(s)
    n, D = X.shape
    out = np.zeros(2 + 4 * D)
    _lib().oracle_moment_match_diag(n, D, _p(X), _p(Ky_inv), _p(beta), _p(lambdas), float(sigma_f),
                                    _p(u), _p(s), _p(out))
    return out[0], out[1], out[2:]


def c_rollout_cost_grad(X, Ky_invs, betas, lambdas, sigma_fs, x0, U, gamma, Q, R, R_delta=None,
                        last_u=None, x_ref=None, u_ref=None):
    """Cost, d cost / d U, means, variances for one control sequence (C oracle)."""
    X = _c(X)
    n, D = X.shape
    Kinv = _c(np.stack(Ky_invs))
    betas = _c(np.stack(betas))
    E = Kinv.shape[0]
    U = _c(U)
    H, m = U.shape
    lambdas = _c(lambdas); sigma_fs = _c(sigma_fs); x0 = _c(x0); Q = _c(Q); R = _c(R)
    Rd = _c(R_delta)
    lu = _c(last_u) if last_u is not None else np.zeros(m)
    xr = _c(x_ref) if x_ref is not None else np.zeros(E)
    ur = _c(u_ref) if u_ref is not None else np.zeros(m)
    cost_ = np.zeros(1); grad = np.zeros((H, m)); means = np.zeros((H + 1, E)); vars_ = np.zeros((H + 1, E))
    _lib().oracle_rollout_cost_grad(n, D, E, m, H, _p(X), _p(Kinv), _p(betas), _p(lambdas), _p(sigma_fs),
                                    _p(x0), _p(U), float(gamma), _p(Q), _p(R), _p(Rd), _p(lu), _p(xr),
                                    _p(ur), _p(cost_), _p(grad), _p(means), _p(vars_))
    return float(cost_[0]), grad, means, vars_
