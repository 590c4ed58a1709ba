"""RiskSensitiveMPC with the reference's Python interface (reference `src/mpc.py:7-330`).

`objective(x)` / `gradient(x)` are the cyipopt callbacks; both are served by ONE fused device call
(rollout + risk-sensitive cost + exact adjoint, `gpmpc_rollout_cost_grad`) whose result is cached on the
bytes of `x` -- a superset of the reference's cache, which returns the gradient of the last objective call
whatever `x` is (`src/mpc.py:245-255`).
"""
from __future__ import annotations

import numpy as np
import torch

from .backend import F64
from .dynamics import Dynamics


def _to_numpy(v):
    if isinstance(v, torch.Tensor):
        return v.detach().cpu().numpy().astype(np.float64)
    return np.asarray(v, dtype=np.float64)


class RiskSensitiveMPC:
    """MPC with the risk-sensitive cost  sum_i [1/gamma log det(I + gamma Q Sigma_i) +
    (x_i-x_ref)^T (Q^-1 + gamma Sigma_i)^-1 (x_i-x_ref)] + input and input-rate terms."""

    def __init__(self, gamma, horizon, state_dim, input_dim, Q, R, R_delta=None):
        self.gamma = gamma
        self.horizon = horizon
        self.state_dim = state_dim
        self.input_dim = input_dim
        self.Q = Q
        self.R = R
        self.R_delta = R_delta
        self.dynamics = Dynamics(self.state_dim, self.input_dim, nominal_models=None)

        self.device = self.dynamics.device
        self.Q_tor = torch.tensor(self.Q, device=self.device).type(F64)
        self.R_tor = torch.tensor(self.R, device=self.device).type(F64)
        self.R_delta_tor = None if R_delta is None else torch.tensor(self.R_delta, device=self.device).type(F64)

        self.x_ref = torch.zeros(self.state_dim, device=self.device)
        self.u_ref = torch.zeros(self.input_dim, device=self.device)

        # state of the current IPOPT iteration (names kept from the reference)
        self.curr_cost = None
        self.curr_grad = None
        self.curr_state = None
        self.backward_taken = False
        self.curr_u = None
        self._cache_key = None

        self.last_traj = np.random.standard_normal(size=(self.horizon * self.input_dim,))
        self.ub = [1e16 for _ in range(self.input_dim)]
        self.lb = [-1e16 for _ in range(self.input_dim)]
        self.train_empty = True
        # extension (off by default = the reference's variance-only rollout): propagate the full state covariance
        self.full_covariance = False
        self.n_evals = 0
        self._host_cache = {}

    def _host(self, name, v):
        """Host fp64 copy of an attribute; device tensors are copied once per (tensor, in-place version) instead
        of once per callback (each copy is a device synchronisation on the IPOPT critical path)."""
        if not isinstance(v, torch.Tensor):
            return np.asarray(v, dtype=np.float64)
        key = (id(v), v._version)
        hit = self._host_cache.get(name)
        if hit is None or hit[0] != key:
            hit = (key, v, _to_numpy(v))                 # keeps `v` alive so that its id cannot be reused
            self._host_cache[name] = hit
        return hit[2]

    # ---- setters (src/mpc.py:72-116) ---------------------------------------------------------------
    def set_ub(self, ub):
        assert len(ub) == self.input_dim
        self.ub = ub

    def set_lb(self, lb):
        assert len(lb) == self.input_dim
        self.lb = lb

    def set_xref(self, x_ref):
        assert len(x_ref) == self.state_dim
        self.x_ref = torch.tensor(x_ref, device=self.device).type(F64)

    def set_uref(self, u_ref):
        assert len(u_ref) == self.input_dim
        self.u_ref = torch.tensor(u_ref, device=self.device).type(F64)

    # ---- cost on explicit trajectories (host-side helpers; general covariances) -----------------
    def cost(self, x, u, sig, x_ref, u_ref):
        """NumPy cost (`src/mpc.py:118-154`; no input-rate term, like the reference)."""
        E = self.state_dim
        Qi = np.linalg.inv(self.Q)
        total = 0
        for i in range(self.horizon + 1):
            e = x[i, :] - x_ref
            total += np.log(np.linalg.det(np.identity(E) + self.gamma * self.Q @ sig[i, :, :])) / self.gamma
            total += e.T @ np.linalg.inv(Qi + self.gamma * sig[i, :, :]) @ e
        for j in range(self.horizon):
            d = u[j, :] - u_ref
            total += d.T @ self.R @ d
        return total

    def cost_torch(self, x, u, sig, x_ref, u_ref):
        """Torch cost on explicit trajectories (`src/mpc.py:156-200`), differentiable."""
        E = self.state_dim
        eye = torch.eye(E, device=self.device)
        Qi = torch.linalg.inv(self.Q_tor)
        total = 0
        for i in range(self.horizon + 1):
            e = x[i] - x_ref
            total = total + torch.log(torch.linalg.det(eye + self.gamma * self.Q_tor @ sig[i])) / self.gamma
            total = total + e @ torch.linalg.inv(Qi + self.gamma * sig[i]) @ e
        for j in range(self.horizon):
            d = u[j, :] - u_ref
            total = total + d @ self.R_tor @ d
        if self.R_delta_tor is not None:
            first = torch.tensor(self.last_traj[0:self.input_dim], device=self.device).type(F64)[None, :]
            du = torch.diff(torch.concatenate((first, u), dim=0), dim=0)
            for j in range(self.horizon):
                total = total + du[j, :] @ self.R_delta_tor @ du[j, :]
        return total

    # ---- cyipopt callbacks (src/mpc.py:202-267) ------------------------------------------------------
    def _evaluate(self, x):
        H, m = self.horizon, self.input_dim
        U = np.ascontiguousarray(np.asarray(x, dtype=np.float64).reshape(1, H, m))
        x0 = self._host("curr_state", self.curr_state).reshape(1, self.state_dim)
        dyn = self.dynamics
        dyn._require_data()
        dyn._sync_propagation_hypers()
        last_u = None
        if self.R_delta is not None:
            last_u = np.asarray(self.last_traj[0:m], dtype=np.float64).reshape(1, m)
        cost, grad, _, _ = dyn._bundle.cost_grad(x0, U, np.array([float(self.gamma)]), self._host("Q", self.Q),
                                                 self._host("R", self.R),
                                                 None if self.R_delta is None else self._host("R_delta", self.R_delta),
                                                 last_u, self._host("x_ref", self.x_ref), self._host("u_ref", self.u_ref),
                                                 full=self.full_covariance)
        dyn._tape_serial += 1
        self.n_evals += 1
        self.curr_cost = float(cost[0])
        self.curr_grad = grad[0]
        self.curr_u = U[0]
        self._cache_key = U.tobytes()

    def objective(self, x):
        """Cost of the action trajectory x (flattened (horizon, input_dim)); a Python float (NaN allowed)."""
        self._evaluate(x)
        self.backward_taken = False
        return self.curr_cost

    def gradient(self, x):
        """d cost / d x as an (horizon, input_dim) array; cyipopt flattens it row-major."""
        key = np.ascontiguousarray(np.asarray(x, dtype=np.float64).reshape(1, self.horizon, self.input_dim)).tobytes()
        if self.curr_cost is None or key != self._cache_key:
            self._evaluate(x)
        self.backward_taken = True
        return self.curr_grad

    def constraints(self, x):
        return 0

    def jacobian(self, x):
        return np.zeros(x.shape)

    # ---- solve (src/mpc.py:269-330) ------------------------------------------------------------------
    def get_optimal_trajectory(self, curr_state):
        """One NLP solve from x = 0 with the reference's IPOPT options.  If `cyipopt` is not importable a
        bounded L-BFGS-B (scipy) drives the same callbacks instead (IPOPT-iterate parity then does not
        apply)."""
        if self.train_empty:
            if self.dynamics.gpr_err[0].num_train > 0:
                self.train_empty = False
            else:
                return np.zeros((self.horizon, self.input_dim))

        self.curr_state = torch.tensor(curr_state, device=self.device).type(F64)
        x0 = np.zeros(shape=len(self.last_traj))
        lb = self.horizon * list(self.lb)
        ub = self.horizon * list(self.ub)
        try:
            import cyipopt
        except ImportError:
            cyipopt = None
        if cyipopt is not None and hasattr(cyipopt, "Problem"):
            nlp = cyipopt.Problem(n=len(x0), m=0, problem_obj=self, lb=lb, ub=ub, cl=[0], cu=[0])
            for key, val in (("mu_strategy", "adaptive"), ("accept_every_trial_step", "yes"), ("max_iter", 300),
                             ("tol", 1e-4), ("acceptable_tol", 1e-4), ("constr_viol_tol", 1e-4),
                             ("compl_inf_tol", 1e-4), ("dual_inf_tol", 1e-4), ("mu_target", 1e-4),
                             ("acceptable_iter", 3), ("sb", "yes"), ("print_level", 0)):
                nlp.add_option(key, val)
            x, _info = nlp.solve(x0)
        else:
            from scipy.optimize import minimize

            last = {"z": x0.copy(), "f": None}

            def fun(z):
                c = self.objective(z)
                g = np.asarray(self.gradient(z), dtype=np.float64).reshape(-1)
                if np.isfinite(c) and np.all(np.isfinite(g)):
                    last["z"], last["f"] = np.array(z, dtype=np.float64), c
                    return c, g
                # A non-finite cost (the NaN of log(det) < 0, `src/mpc.py:183`) is an evaluation error: IPOPT shortens
                # the step.  L-BFGS-B's line search cannot take NaN, so it is shown a steep bowl centred on the last
                # finite iterate instead, which makes it back off towards that iterate in the same way.
                d = np.asarray(z, dtype=np.float64) - last["z"]
                scale = 1e3 * max(1.0, abs(last["f"] or 0.0))
                return (last["f"] or 0.0) + scale * (1.0 + d @ d), 2.0 * scale * d
            bounds = [(None if l <= -1e15 else l, None if u >= 1e15 else u) for l, u in zip(lb, ub)]
            res = minimize(fun, x0, jac=True, method="L-BFGS-B", bounds=bounds,
                           options={"maxiter": 300, "ftol": 1e-10, "gtol": 1e-4})
            x = res.x
        self.last_traj = x
        return np.reshape(x, (self.horizon, self.input_dim))
