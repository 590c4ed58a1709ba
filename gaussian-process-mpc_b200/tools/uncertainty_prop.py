"""Moment matching free functions with the reference's signatures
(reference `src/tools/uncertainty_prop.py:296-465`): exact mean / variance / cross-covariance of GP outputs
for a Gaussian input N(u, S) with a FULL covariance S, evaluated by libgpmpc's generic pair-sum kernels
(`gpmpc_moment_match_raw`, `gpmpc_covariance_raw`).

These are the stateless forms: every call ships the caller's tensors to the library.  The rollout does not
go through them -- it uses the fitted bundle and the batched kernels (`Dynamics.forward_propagate_torch`).
Outputs are detached device tensors (the reference's tests use values only; gradients w.r.t. the control
sequence are provided by the rollout adjoint instead).
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


def _scalar(x):
    return float(x.item()) if isinstance(x, torch.Tensor) else float(x)


def mean_prop_torch(Ky_inv, lambdas, u, S, X_train, y_train, sigma_f=1):
    """Mean of the GP output; returns (0-D tensor, {'beta': [n], 'l': [n]})."""
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
