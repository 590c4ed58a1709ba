"""Moment matching free functions with the reference's signatures
(reference `src/tools/uncertainty_prop.py:296-465`): exact mean / variance / cross-covariance of GP outputs
for a Gaussian input N(u, S) with a FULL covariance S, evaluated by libgpmpc's generic pair-sum kernels
(`gpmpc_moment_match_raw`, `gpmpc_covariance_raw`).

These are the stateless forms: every call ships the caller's tensors to the library.  The rollout does not
go through them -- it uses the fitted bundle and the batched kernels (`Dynamics.forward_propagate_torch`).
Outputs are detached device tensors (the reference's tests use values only; gradients w.r.t. the control
sequence are provided by the rollout adjoint instead).  The reference's functions are differentiable torch code;
to keep a caller that composes them the reference way from silently training on missing gradients, they raise
when u or S asks for a gradient.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from ..backend import F64, _ptr, as_f64, default_bundle
from .._lib import check


def _dev(t):
    if isinstance(t, torch.Tensor) and t.is_cuda:
        return t.device
    return torch.device("cuda", torch.cuda.current_device())


def _no_autograd(*tensors):
    for t in tensors:
        if isinstance(t, torch.Tensor) and t.requires_grad and torch.is_grad_enabled():
            raise RuntimeError("the moment-matching free functions return values only; differentiate through "
                               "Dynamics.forward_propagate_torch (device adjoint) or detach the inputs")


def _scalar(x):
    return float(x.item()) if isinstance(x, torch.Tensor) else float(x)


def mean_prop_torch(Ky_inv, lambdas, u, S, X_train, y_train, sigma_f=1):
    """Mean of the GP output; returns (0-D tensor, {'beta': [n], 'l': [n]})."""
    _no_autograd(u, S)
    dev = _dev(Ky_inv)
    b = default_bundle(dev.index or 0)
    b._sync_stream()
    Kinv = as_f64(Ky_inv, dev); X = as_f64(X_train, dev)
    n, D = X.shape
    y = as_f64(y_train, dev).reshape(-1)
    lam = as_f64(lambdas, dev); uu = as_f64(u, dev); SS = as_f64(S, dev)
    mean = torch.empty(1, dtype=F64, device=dev)
    beta = torch.empty(n, dtype=F64, device=dev)
    l = torch.empty(n, dtype=F64, device=dev)
    check(b.h, b.lib.gpmpc_moment_match_raw(b.h, n, D, _ptr(Kinv), _ptr(lam), _ptr(uu), _ptr(SS), _ptr(X), _ptr(y),
                                            _scalar(sigma_f), _ptr(mean), None, _ptr(beta), _ptr(l)),
          "gpmpc_moment_match_raw")
    return mean[0], {'beta': beta, 'l': l}


def variance_prop_torch(Ky_inv, lambdas, u, S, X_train, mean, beta, sigma_f=1):
    """Variance of the GP output (latent; no noise term) given the mean and beta of mean_prop_torch."""
    _no_autograd(u, S)
    dev = _dev(Ky_inv)
    b = default_bundle(dev.index or 0)
    b._sync_stream()
    Kinv = as_f64(Ky_inv, dev); X = as_f64(X_train, dev)
    n, D = X.shape
    lam = as_f64(lambdas, dev); uu = as_f64(u, dev); SS = as_f64(S, dev)
    m = torch.tensor([_scalar(mean)], dtype=F64, device=dev)
    bt = as_f64(beta, dev).reshape(-1)
    var = torch.empty(1, dtype=F64, device=dev)
    check(b.h, b.lib.gpmpc_moment_match_raw(b.h, n, D, _ptr(Kinv), _ptr(lam), _ptr(uu), _ptr(SS), _ptr(X), None,
                                            _scalar(sigma_f), _ptr(m), _ptr(var), _ptr(bt), None),
          "gpmpc_moment_match_raw")
    return var[0]


def covariance_prop_torch(lambdas1, lambdas2, u, S, X_train, mean1, mean2, beta1, beta2, sigma_f1=1, sigma_f2=1,
                          bugcompat=False):
    """Cross-covariance of two GP outputs.

    Evaluates the published formula (and the reference's NumPy twin, `uncertainty_prop.py:187-236`).  The
    reference's torch function transposes the cross term (`:446`), which only matters when lambdas1 is not
    proportional to lambdas2; pass bugcompat=True to reproduce that function bit-for-formula."""
    _no_autograd(u, S)
    dev = _dev(X_train)
    b = default_bundle(dev.index or 0)
    b._sync_stream()
    X = as_f64(X_train, dev)
    n, D = X.shape
    l1 = as_f64(lambdas1, dev); l2 = as_f64(lambdas2, dev); uu = as_f64(u, dev); SS = as_f64(S, dev)
    b1 = as_f64(beta1, dev).reshape(-1); b2 = as_f64(beta2, dev).reshape(-1)
    cov = torch.empty(1, dtype=F64, device=dev)
    check(b.h, b.lib.gpmpc_covariance_raw(b.h, n, D, _ptr(l1), _ptr(l2), _ptr(uu), _ptr(SS), _ptr(X), _scalar(mean1),
                                          _scalar(mean2), _ptr(b1), _ptr(b2), _scalar(sigma_f1), _scalar(sigma_f2),
                                          int(bool(bugcompat)), _ptr(cov)), "gpmpc_covariance_raw")
    return cov[0]


# ---- NumPy-interface twins (reference `src/tools/uncertainty_prop.py:6-44,91-136,187-236`) -------------------------
# The reference's versions are O(n^2) Python loops kept as test oracles (sigma_f = 1, they take the evidence matrix
# K = Ky rather than its inverse).  Same signatures and return types here, evaluated by the same device kernels as
# the torch functions above.
def _inv_dev(K, dev):
    return torch.linalg.inv(torch.as_tensor(np.asarray(K, dtype=np.float64), device=dev))


def mean_prop(K, Lambda, u, S, X_train, y_train):
    """Mean of the GP output for x ~ N(u, S); returns (float, {'beta': ndarray, 'l': ndarray})."""
    dev = torch.device("cuda", torch.cuda.current_device())
    mean, params = mean_prop_torch(_inv_dev(K, dev), np.diag(np.asarray(Lambda, dtype=np.float64)).copy(), u, S, X_train,
                                   y_train, 1.0)
    return float(mean.item()), {'beta': params['beta'].cpu().numpy(), 'l': params['l'].cpu().numpy()}


def variance_prop(K, Lambda, u, S, X_train, y_train):
    """Variance of the GP output (sigma_f = 1): 1 - tr((K^-1 - beta beta^T) L) - mean^2."""
    dev = torch.device("cuda", torch.cuda.current_device())
    Kinv = _inv_dev(K, dev)
    lam = np.diag(np.asarray(Lambda, dtype=np.float64)).copy()
    mean, params = mean_prop_torch(Kinv, lam, u, S, X_train, y_train, 1.0)
    return float(variance_prop_torch(Kinv, lam, u, S, X_train, mean, params['beta'], 1.0).item())


def covariance_prop(K1, K2, Lambda1, Lambda2, u, S, X_train, y_train):
    """Covariance of two GP outputs that share the targets y_train (as the reference's signature has it)."""
    dev = torch.device("cuda", torch.cuda.current_device())
    lam1 = np.diag(np.asarray(Lambda1, dtype=np.float64)).copy(); lam2 = np.diag(np.asarray(Lambda2, dtype=np.float64)).copy()
    m1, p1 = mean_prop_torch(_inv_dev(K1, dev), lam1, u, S, X_train, y_train, 1.0)
    m2, p2 = mean_prop_torch(_inv_dev(K2, dev), lam2, u, S, X_train, y_train, 1.0)
    return float(covariance_prop_torch(lam1, lam2, u, S, X_train, m1, m2, p1['beta'], p2['beta'], 1.0, 1.0).item())


# ---- Monte-Carlo checkers (reference `src/tools/uncertainty_prop.py:47-88,139-184,239-292`) ---------------------------
# Test utilities in the reference (T = 10 000 samples, Python loops over samples and training points); same signatures
# and estimators here, vectorised with torch on the device (plain library ops: they are checkers, not on the hot path).
_MC_SAMPLES = 10000


def _mc_posterior(K, Lambda, X_star, X_train, y_train, sigma_f, dev):
    """Posterior mean [T] and variance [T] of the GP at the sampled inputs (`:76-86`, `:170-182`)."""
    lam = torch.as_tensor(np.diag(np.asarray(Lambda, dtype=np.float64)).copy(), device=dev)
    X = torch.as_tensor(np.asarray(X_train, dtype=np.float64), device=dev)
    y = torch.as_tensor(np.asarray(y_train, dtype=np.float64).reshape(-1), device=dev)
    Kinv = _inv_dev(K, dev)
    d2 = torch.cdist(X_star / torch.sqrt(lam), X / torch.sqrt(lam)).square()
    kv = sigma_f ** 2 * torch.exp(-0.5 * d2)                       # [T, n]
    mu = kv @ (Kinv @ y)
    sig2 = sigma_f ** 2 - ((kv @ Kinv) * kv).sum(dim=1)
    return mu, sig2


def _mc_inputs(u, S, dev, samples):
    xs = np.random.multivariate_normal(np.asarray(u, dtype=np.float64), np.asarray(S, dtype=np.float64), size=samples)
    return torch.as_tensor(xs, device=dev)


def mean_prop_mc(K, Lambda, u, S, X_train, y_train, sigma_f=1):
    """Monte-Carlo estimate of the mean of the GP output for x ~ N(u, S)."""
    dev = torch.device("cuda", torch.cuda.current_device())
    mu, _ = _mc_posterior(K, Lambda, _mc_inputs(u, S, dev, _MC_SAMPLES), X_train, y_train, sigma_f, dev)
    return float(mu.mean().item())


def variance_prop_mc(K, Lambda, u, S, X_train, y_train, sigma_f=1):
    """Monte-Carlo estimate of the variance: E[sigma^2(x)] + Var[mu(x)] (law of total variance, `:184`)."""
    dev = torch.device("cuda", torch.cuda.current_device())
    mu, sig2 = _mc_posterior(K, Lambda, _mc_inputs(u, S, dev, _MC_SAMPLES), X_train, y_train, sigma_f, dev)
    return float(sig2.mean().item() + mu.var(unbiased=False).item())


def covariance_prop_mc(K1, K2, Lambda1, Lambda2, u, S, X_train, y_train, sigma_f1=1, sigma_f2=1):
    """Monte-Carlo estimate of the covariance of two GP outputs: sample x, then f1, f2 independently given x (`:276-292`)."""
    dev = torch.device("cuda", torch.cuda.current_device())
    xs = _mc_inputs(u, S, dev, _MC_SAMPLES)
    mu1, s1 = _mc_posterior(K1, Lambda1, xs, X_train, y_train, sigma_f1, dev)
    mu2, s2 = _mc_posterior(K2, Lambda2, xs, X_train, y_train, sigma_f2, dev)
    f1 = mu1.cpu().numpy() + np.sqrt(np.maximum(s1.cpu().numpy(), 0.0)) * np.random.standard_normal(_MC_SAMPLES)
    f2 = mu2.cpu().numpy() + np.sqrt(np.maximum(s2.cpu().numpy(), 0.0)) * np.random.standard_normal(_MC_SAMPLES)
    return float(np.cov(f1, f2)[0, 1])
