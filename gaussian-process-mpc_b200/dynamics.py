"""Dynamics with the reference's Python interface (reference `src/dynamics.py:8-191`): a bundle of
`state_dim` GPs that share the training inputs [state, action] and predict one next-state coordinate each.

The whole bundle lives in ONE libgpmpc handle; `forward_propagate_torch` is a single fused rollout on the
device with an exact hand-written adjoint exposed to autograd.
"""
from __future__ import annotations

import numpy as np
import torch

from .backend import GPBundle, F64
from .gpr import GaussianProcessRegression


class _Rollout(torch.autograd.Function):
    """means, vars = rollout(x0, actions) with d/d actions and d/d x0 from gpmpc_rollout_vjp."""

    @staticmethod
    def forward(ctx, dyn, x0, actions):
        means, vars_ = dyn._bundle.rollout(x0[None, :], actions[None, :, :])
        dyn._tape_serial += 1
        ctx.dyn = dyn
        ctx.serial = dyn._tape_serial
        ctx.save_for_backward(x0, actions)
        return means[0], vars_[0]

    @staticmethod
    def backward(ctx, gmeans, gvars):
        dyn = ctx.dyn
        x0, actions = ctx.saved_tensors
        H = actions.shape[0]
        if ctx.serial != dyn._tape_serial:           # another rollout overwrote the tape: replay forward
            dyn._bundle.rollout(x0[None, :], actions[None, :, :])
            dyn._tape_serial += 1
            ctx.serial = dyn._tape_serial
        gU, gx0 = dyn._bundle.rollout_vjp(1, H, gmeans.contiguous()[None], gvars.contiguous()[None], want_gx0=True)
        return None, gx0[0], gU[0]


class _RolloutFull(torch.autograd.Function):
    """means, covs = full-covariance rollout(x0, actions) with d/d actions and d/d x0 from gpmpc_rollout_full_vjp."""

    @staticmethod
    def forward(ctx, dyn, x0, actions):
        means, covs = dyn._bundle.rollout_full(x0[None, :], actions[None, :, :])
        dyn._tape_serial += 1
        ctx.dyn = dyn
        ctx.serial = dyn._tape_serial
        ctx.save_for_backward(x0, actions)
        return means[0], covs[0]

    @staticmethod
    def backward(ctx, gmeans, gcovs):
        dyn = ctx.dyn
        x0, actions = ctx.saved_tensors
        H = actions.shape[0]
        if ctx.serial != dyn._tape_serial:           # another rollout overwrote the tape: replay forward
            dyn._bundle.rollout_full(x0[None, :], actions[None, :, :])
            dyn._tape_serial += 1
            ctx.serial = dyn._tape_serial
        gU, gx0 = dyn._bundle.rollout_full_vjp(1, H, gmeans.contiguous()[None], gcovs.contiguous()[None], want_gx0=True)
        return None, gx0[0], gU[0]


class Dynamics(object):

    def __init__(self, state_dim, action_dim, nominal_models=None):
        self.state_dim = state_dim
        self.action_dim = action_dim
        self.nominal_models = nominal_models
        D = state_dim + action_dim
        self.gpr_err = [GaussianProcessRegression(D, None if nominal_models is None else nominal_models[i],
                                                  _owner=self, _index=i) for i in range(state_dim)]
        self.device = self.gpr_err[0].device
        self._bundle = None
        self._X = None          # [n, D] device tensor shared by all members
        self._Y = None          # [n, E]
        self._prop_key = None   # hyper-parameter snapshot last pushed to the device
        self._fit_key = None    # hyper-parameter snapshot of the last full fit
        self.incremental = True # use gpmpc_append_point for single observations
        self._tape_serial = 0

    # ---- data ingestion (src/dynamics.py:39-60) ----------------------------------------------
    @staticmethod
    def _stack_observations(state, action, next_state, state_dim, action_dim):
        """Host-side layout logic: one (1-D) or many (2-D) observations -> X [k, D], Y [k, E]."""
        state = np.asarray(state, dtype=np.float64)
        action = np.asarray(action, dtype=np.float64)
        next_state = np.asarray(next_state, dtype=np.float64)
        if state.ndim == 1:
            x = np.concatenate((state, action.reshape(-1)))[None, :]
            y = next_state.reshape(1, state_dim)
        else:
            if action.ndim == 1:
                action = action[:, None]
            x = np.concatenate((state, action), axis=1)
            y = next_state.reshape(-1, state_dim)
        assert x.shape[1] == state_dim + action_dim, "state/action dimensions do not match the model"
        return x, y

    def append_train_data(self, state, action, next_state):
        x, y = self._stack_observations(state, action, next_state, self.state_dim, self.action_dim)
        x = torch.tensor(x).type(F64).to(self.device)
        y = torch.tensor(y).type(F64).to(self.device)
        if self._X is None:
            self._X, self._Y = x, y
        else:
            self._X = torch.cat((self._X, x), dim=0)
            self._Y = torch.cat((self._Y, y), dim=0)
        # one new observation and unchanged hyper-parameters: bordered O(n^2) update (the closed loop of
        # src/simulator.py:55 appends one point per step); otherwise, or when the library asks for it, refit
        if (self.incremental and self._bundle is not None and x.shape[0] == 1
                and self._fit_key == tuple(g._hyper_key() for g in self.gpr_err)
                and self._bundle.append_point(x[0], y[0])):
            for g in self.gpr_err:
                g.num_train = self._X.shape[0]
                g._mats = {}
            return
        self._fit_all()

    def _collect_hypers(self):
        lam = np.stack([g.get_lambdas().astype(np.float64) for g in self.gpr_err])
        sf = np.array([g.get_sigma_f() for g in self.gpr_err], dtype=np.float64)
        nv = np.array([g._hyper_values()[2] for g in self.gpr_err], dtype=np.float64)
        return lam, sf, nv

    def _fit_all(self):
        if self._bundle is None:
            self._bundle = GPBundle(self.state_dim + self.action_dim, self.state_dim, self.device.index or 0)
        lam, sf, nv = self._collect_hypers()
        self._bundle.fit(self._X, self._Y, lam, sf, nv)
        self._prop_key = tuple(g._hyper_key() for g in self.gpr_err)
        self._fit_key = self._prop_key
        for g in self.gpr_err:
            g.num_train = self._X.shape[0]
            g._mats = {}

    def _refit_member(self, index, lam, sf, nv):
        self._bundle.refit_output(index, None, lam, sf, nv)
        self._prop_key = None            # forces a re-sync of the other members' propagation hypers
        self._fit_key = None             # mixed fit state: the next append does a full fit

    def _sync_propagation_hypers(self):
        """The reference reads lambdas / sigma_f at rollout time (src/dynamics.py:171,173) while Ky_inv
        stays from the last build; push changed values (no device sync unless something changed)."""
        key = tuple(g._hyper_key() for g in self.gpr_err)
        if key != self._prop_key:
            lam, sf, _ = self._collect_hypers()
            self._bundle.set_propagation_hypers(lam, sf)
            self._prop_key = key

    def _require_data(self):
        if self._bundle is None or self._X is None:
            raise RuntimeError("no training data: call append_train_data first")

    # ---- rollouts -------------------------------------------------------------------------------
    def forward_propagate_torch(self, horizon, curr_state, actions, full=False):
        """Variance-only moment-matched rollout (`src/dynamics.py:126-191`).

        Returns (list of H+1 mean tensors [E], list of H+1 covariance tensors [E,E]); autograd flows from
        `actions` (and `curr_state`) through the device adjoint.  full=True keeps the cross-covariances between the
        outputs (the reference's TODO at `src/dynamics.py:184`): Sigma_t is then a full matrix."""
        self._require_data()
        self._sync_propagation_hypers()
        x0 = curr_state.to(self.device).type(F64)
        U = actions.to(self.device).type(F64)[:horizon]
        if full:
            means, covs = _RolloutFull.apply(self, x0, U)
            return ([curr_state] + [means[t] for t in range(1, horizon + 1)],
                    [covs[t] for t in range(horizon + 1)])
        means, vars_ = _Rollout.apply(self, x0, U)
        state_means = [curr_state] + [means[t] for t in range(1, horizon + 1)]
        state_covars = [1e-3 * torch.eye(self.state_dim, device=self.device).type(F64)]
        state_covars += [torch.diag(vars_[t]) for t in range(1, horizon + 1)]
        return state_means, state_covars

    def forward_propagate(self, horizon, curr_state, actions):
        """NumPy-interface rollout (`src/dynamics.py:62-124`): arrays (H+1, E) and (H+1, E, E).

        The reference's NumPy twin is an O(n^2) Python loop kept as a test oracle; here it is the same
        device rollout as the torch method (action variance fp32(1e-3) as in the torch method)."""
        self._require_data()
        self._sync_propagation_hypers()
        x0 = np.asarray(curr_state, dtype=np.float64).reshape(1, self.state_dim)
        U = np.asarray(actions, dtype=np.float64).reshape(1, -1, self.action_dim)[:, :horizon]
        means, vars_ = self._bundle.rollout(x0, U, out_device=False)
        self._tape_serial += 1
        covs = np.zeros((horizon + 1, self.state_dim, self.state_dim))
        for t in range(horizon + 1):
            covs[t] = np.diag(vars_[0, t])
        return means[0], covs

    def forward_propagate_full(self, horizon, curr_state, actions):
        """Full-covariance moment-matched rollout (the reference's TODO at `src/dynamics.py:184`):
        Sigma_t keeps the cross-covariances between the outputs, computed with the published formula
        (`src/tools/uncertainty_prop.py:187-236`); the input covariance is blockdiag(Sigma_{t-1}, fp32(1e-3) I) as in
        the variance-only rollout.  NumPy arrays (H+1, E), (H+1, E, E); one batched device rollout."""
        self._require_data()
        self._sync_propagation_hypers()
        x0 = np.asarray(curr_state, dtype=np.float64).reshape(1, self.state_dim)
        U = np.asarray(actions, dtype=np.float64).reshape(1, -1, self.action_dim)[:, :horizon]
        means, covs = self._bundle.rollout_full(x0, U, out_device=False)
        self._tape_serial += 1
        return means[0], covs[0]
