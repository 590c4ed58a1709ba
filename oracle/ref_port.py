"""Torch-CPU fp64 port of the reference's hot path with the reference's own operation sequence --
TEST / BASELINE INFRASTRUCTURE ONLY (never imported by the product package).

Why a second oracle: `oracle/oracle.py` + `oracle/gpmpc_oracle.c` restate the *mathematics* as O(n^2)
pair sums.  This file restates the *cost structure* of the reference: explicit LU inverse, beta recomputed
by a gemv each step, ~12 dense [n,n] temporaries, the n^3 `mm` whose trace is taken
(`src/tools/uncertainty_prop.py:399`) and reverse-mode autograd for the gradient (`src/mpc.py:251`).  It is
what `bench.py --impl reference` and `cpu_baseline` time on the GPU box's host cores, because the Python
reference itself cannot travel to that box.

Parity status: PINNED against golden vectors of the unmodified reference (tests/test_oracle.py).
Citations are relative to /root/reference.
"""
from __future__ import annotations

import numpy as np
import torch

F64 = torch.float64


def build_inverse(X, lambdas, sigma_f, sigma_n):
    """Kf, Ky, Ky_inv -- `src/gpr.py:159-171`."""
    Xs = X * torch.sqrt(1.0 / lambdas)
    d = torch.cdist(Xs, Xs, p=2)
    Kf = sigma_f ** 2 * torch.exp(-0.5 * d.square())
    Ky = Kf + sigma_n ** 2 * torch.eye(X.shape[0], dtype=torch.float32)
    return Kf, Ky, torch.linalg.inv(Ky)


def mm_mean(Ky_inv, lambdas, u, S, X, y, sigma_f):
    """`src/tools/uncertainty_prop.py:296-338`."""
    beta = Ky_inv @ y
    D = S.shape[0]
    B = torch.linalg.inv(S + torch.diag(lambdas))
    V = u - X
    quad = ((V @ B) * V).sum(dim=1)
    det = torch.linalg.det(torch.diag(1.0 / lambdas) @ S + torch.eye(D))
    l = det ** (-0.5) * torch.exp(-0.5 * quad) * sigma_f ** 2
    return torch.dot(beta, l), beta


def mm_variance(Ky_inv, lambdas, u, S, X, mean, beta, sigma_f):
    """`src/tools/uncertainty_prop.py:341-399`, including the dense mm + trace at :399."""
    D = S.shape[0]
    A = torch.linalg.inv(torch.diag(lambdas) / 2.0 + S)
    det = torch.linalg.det(2.0 * torch.diag(1.0 / lambdas) @ S + torch.eye(D)) ** (-0.5)
    uAX = (u @ A @ X.mT)[:, None]
    q = u @ A @ u + X @ A @ X.mT - uAX - uAX.mT
    qd = torch.diag(q)[:, None]
    a_part = torch.exp(-0.125 * (qd + 2.0 * q + qd.mT))
    Xs = X * torch.sqrt(1.0 / lambdas)
    lam_part = torch.exp(-0.25 * torch.cdist(Xs, Xs, p=2).square())
    L = det * a_part * lam_part * sigma_f ** 4
    return sigma_f ** 2 - torch.trace((Ky_inv - torch.outer(beta, beta)) @ L) - mean ** 2


def propagate(X, Ky_invs, Ys, lambdas, sigma_fs, x0, U):
    """`src/dynamics.py:126-191`.  Returns lists of means [E] and covariances [E,E]."""
    E = len(Ky_invs)
    m = U.shape[1]
    means = [x0]
    covs = [1e-3 * torch.eye(E).type(F64)]
    for t in range(1, U.shape[0] + 1):
        u = torch.cat((means[t - 1], U[t - 1, :]))
        top = torch.cat((covs[t - 1], torch.zeros((E, m))), dim=1)
        bot = torch.cat((torch.zeros((m, E)), 1e-3 * torch.eye(m)), dim=1)   # fp32 eye, dynamics.py:162
        S = torch.cat((top, bot), dim=0)
        mu_t, var_t = [], []
        for a in range(E):
            mu, beta = mm_mean(Ky_invs[a], lambdas[a], u, S, X, Ys[:, a], sigma_fs[a])
            mu_t.append(mu)
            var_t.append(mm_variance(Ky_invs[a], lambdas[a], u, S, X, mu, beta, sigma_fs[a]))
        means.append(torch.stack(mu_t))
        covs.append(torch.diag(torch.stack(var_t)))
    return means, covs


def risk_cost(means, U, covs, x_ref, u_ref, gamma, Q, R, R_delta=None, last_u=None):
    """`src/mpc.py:156-200`."""
    E = Q.shape[0]
    Qi = torch.linalg.inv(Q)
    c = 0
    for i in range(len(means)):
        c = c + (1.0 / gamma) * torch.log(torch.linalg.det(torch.eye(E) + gamma * Q @ covs[i]))
        e = means[i] - x_ref
        c = c + e @ torch.linalg.inv(Qi + gamma * covs[i]) @ e
    for j in range(U.shape[0]):
        du = U[j, :] - u_ref
        c = c + du @ R @ du
    if R_delta is not None:
        dU = torch.diff(torch.cat((last_u[None, :], U), dim=0), dim=0)
        for j in range(U.shape[0]):
            c = c + dU[j, :] @ R_delta @ dU[j, :]
    return c


class RefPortProblem:
    """Holds a fitted GP bundle (explicit inverses) and evaluates objective + gradient the way the
    reference's cyipopt callbacks do (`src/mpc.py:202-255`)."""

    def __init__(self, X, Y, lambdas, sigma_fs, sigma_ns, gamma, Q, R, R_delta=None, last_u=None,
                 x_ref=None, u_ref=None, device="cpu"):
        # device="cuda" runs the same eager torch operation sequence on the GPU: what a user of the reference gets on
        # a GPU box (the reference picks cuda:0 when available, `src/gpr.py:22`); used for context by bench.py only
        self.device = torch.device(device)
        t = lambda a: torch.as_tensor(np.asarray(a)).type(F64).to(self.device)      # noqa: E731
        with torch.device(self.device):
            self.X = t(X)
            self.Y = t(Y)
            self.E = self.Y.shape[1]
            self.lambdas = [t(lambdas[a]) for a in range(self.E)]
            self.sigma_fs = [float(s) for s in sigma_fs]
            self.Ky_invs = [build_inverse(self.X, self.lambdas[a], float(sigma_fs[a]), float(sigma_ns[a]))[2]
                            for a in range(self.E)]
            self.gamma = float(gamma)
            self.Q = t(Q)
            self.R = t(R)
            self.Rd = None if R_delta is None else t(R_delta)
            m = self.R.shape[0]
            self.last_u = torch.zeros(m).type(F64) if last_u is None else t(last_u)
            self.x_ref = torch.zeros(self.E).type(F64) if x_ref is None else t(x_ref)
            self.u_ref = torch.zeros(m).type(F64) if u_ref is None else t(u_ref)

    def cost_and_grad(self, x0, U):
        with torch.device(self.device):
            x0 = torch.as_tensor(np.asarray(x0)).type(F64).to(self.device)
            U = torch.as_tensor(np.asarray(U)).type(F64).to(self.device).clone().requires_grad_(True)
            means, covs = propagate(self.X, self.Ky_invs, self.Y, self.lambdas, self.sigma_fs, x0, U)
            c = risk_cost(means, U, covs, self.x_ref, self.u_ref, self.gamma, self.Q, self.R, self.Rd, self.last_u)
            c.backward()
            return c.item(), U.grad.cpu().numpy().copy(), torch.stack(means).detach().cpu().numpy(), \
                torch.stack([torch.diag(s) for s in covs]).detach().cpu().numpy()
