"""GaussianProcessRegression with the reference's Python interface (reference `src/gpr.py:5-370`),
backed by libgpmpc.so.

Same constructor, attributes and method names; the matrices `Kf`, `Ky`, `Ky_inv` are materialised from
the device on first access.  A GPR is either standalone (its own 1-output bundle) or a member of a
`Dynamics` bundle (all members share `X_train`, `src/dynamics.py:33-60`).

Reference quirks that are reproduced on purpose because they move results by more than fp64 rounding:
  * the setters build their tensors from Python floats, i.e. in fp32, before converting to fp64
    (`src/gpr.py:59,72,85`) -- the same torch expression is used here;
  * the noise term is added through an fp32 identity (`src/gpr.py:170`), so the diagonal gets
    float32(sigma_n^2).
Hyper-parameter training: `compute_marginal_likelihood` evaluates the log marginal
likelihood on the device from the Cholesky factor (2 sum log diag L instead of the reference's
log(det(Ky)), which under/overflows for large n, `src/gpr.py:245-247`) and carries an exact analytic
gradient into autograd; `update_hyperparams` is the reference's Adam loop on top of it
(`src/gpr.py:334-370`).  `kernel_matrix_gradient`, `marginal_likelihood_grad` (marked wrong in the reference,
`src/gpr.py:204`) and `update_Ky_inv_mat` (marked "don't use", `src/gpr.py:139`) are not provided.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from .backend import GPBundle, F64


def current_device():
    if torch.cuda.is_available():
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def noise_variance(sigma_n: float) -> float:
    """sigma_n^2 as the reference adds it to the diagonal: through an fp32 eye (`src/gpr.py:170`)."""
    return float(np.float32(float(sigma_n) ** 2))


class _MarginalLikelihood(torch.autograd.Function):
    """ml(log_lambdas, log_sigma_f, log_sigma_n) for the state of the last build; the backward pass hands the
    device-computed gradient 1/2 tr((alpha alpha^T - Ky^-1) dKy/dtheta) to autograd."""

    @staticmethod
    def forward(ctx, gpr, log_lambdas, log_sigma_f, log_sigma_n):
        bundle, a = gpr._bundle_and_index()
        resid = None
        if gpr.f_nom is not None:
            resid = (gpr.y_train - gpr.f_nom(gpr.X_train)).reshape(-1)
        ml, grad = bundle.marginal_likelihood(a, resid, want_grad=True)
        ctx.grad = torch.tensor(grad, dtype=F64, device=log_lambdas.device)
        return torch.full((1, 1), ml, dtype=F64, device=log_lambdas.device)

    @staticmethod
    def backward(ctx, g):
        g = g.reshape(())
        D = ctx.grad.shape[0] - 2
        return None, g * ctx.grad[:D], g * ctx.grad[D], g * ctx.grad[D + 1]


class GaussianProcessRegression(object):
    """Exact GP regression with the squared-exponential ARD kernel (lambdas are squared length-scales)."""

    def __init__(self, x_dim, nominal_model=None, _owner=None, _index=0):
        # the reference hard-codes cuda:0 (`src/gpr.py:22`); one process per GPU uses its current device
        self.device = current_device()
        self.x_dim = x_dim
        self.num_train = 0
        self._X = None                       # (n, x_dim) device tensor (standalone mode)
        self._y = None                       # (n, 1)
        self.log_lambdas = torch.zeros(x_dim, device=self.device).type(F64).requires_grad_()
        self.log_sigma_n = torch.tensor(0.0, device=self.device).type(F64).requires_grad_()
        self.log_sigma_f = torch.tensor(0.0, device=self.device).type(F64).requires_grad_()
        self.f_nom = nominal_model
        self._owner = _owner                 # Dynamics that owns the shared bundle, or None
        self._index = _index
        self._bundle = None                  # standalone 1-output bundle, created at first fit
        self._mats = {}                      # materialised Kf / Ky / Ky_inv
        # like the reference, the optimiser is bound to the tensors created here (`src/gpr.py:46-49`): after a
        # set_* call it keeps training the originals
        self.optimizer = torch.optim.Adam(params=[self.log_lambdas, self.log_sigma_n, self.log_sigma_f],
                                          lr=0.1, betas=(0.9, 0.999), maximize=True)

    # ---- hyper-parameters (src/gpr.py:51-88) ------------------------------------------------
    def set_lambdas(self, lambdas):
        """Like the reference, does not rebuild the matrices; call build_Ky_inv_mat() for that."""
        self.log_lambdas = torch.log(torch.tensor(lambdas, device=self.device)).type(F64).requires_grad_()

    def get_lambdas(self):
        return torch.exp(self.log_lambdas).cpu().detach().numpy()

    def set_sigma_f(self, sigma_f):
        self.log_sigma_f = torch.log(torch.tensor(sigma_f, device=self.device)).type(F64).requires_grad_()

    def get_sigma_f(self):
        return torch.exp(self.log_sigma_f).item()

    def set_sigma_n(self, sigma_n):
        self.log_sigma_n = torch.log(torch.tensor(sigma_n, device=self.device)).type(F64).requires_grad_()

    def get_sigma_n(self):
        return torch.exp(self.log_sigma_n).item()

    _HYPER_ATTRS = ("log_lambdas", "log_sigma_f", "log_sigma_n")

    def __setattr__(self, name, value):
        # every (re)assignment of a hyper-parameter tensor -- by a set_* method or directly -- bumps a generation
        # counter; id() alone cannot tell two tensors apart once the first has been freed and its id recycled
        if name in GaussianProcessRegression._HYPER_ATTRS:
            object.__setattr__(self, "_hyper_gen", getattr(self, "_hyper_gen", 0) + 1)
        object.__setattr__(self, name, value)

    def _hyper_key(self):
        """Cheap change detector (no device sync): assignment generation + in-place version of the three tensors."""
        return (self._hyper_gen,) + tuple(t._version for t in (self.log_lambdas, self.log_sigma_f, self.log_sigma_n))

    def _call_time_hypers(self):
        """[lambdas, sigma_f, sigma_n^2] as the object holds them NOW: the reference evaluates K(X*,X) and K(X*,X*) with
        the current hyper-parameters and adds the fp64 sigma_n^2 to target covariances (`src/gpr.py:268-276,317-329`),
        while Ky^-1 stays from the last build."""
        return np.concatenate([self.get_lambdas().astype(np.float64), [float(self.get_sigma_f())],
                               [float(self.get_sigma_n()) ** 2]])

    def _hyper_values(self):
        return self.get_lambdas().astype(np.float64), float(self.get_sigma_f()), noise_variance(self.get_sigma_n())

    # ---- data ---------------------------------------------------------------------------------
    @property
    def X_train(self):
        return self._owner._X if self._owner is not None else self._X

    @property
    def y_train(self):
        if self._owner is not None:
            Y = self._owner._Y
            return None if Y is None else Y[:, self._index:self._index + 1]
        return self._y

    def _bundle_and_index(self):
        if self._owner is not None:
            return self._owner._bundle, self._index
        return self._bundle, 0

    def append_train_data(self, x, y):
        """Append observations and refit (`src/gpr.py:90-122`).  x: (x_dim,) or (k, x_dim); y: scalar or (k,)."""
        if self._owner is not None:
            raise RuntimeError("this GPR shares its training inputs with a Dynamics bundle; "
                               "use Dynamics.append_train_data")
        if not np.isscalar(y):
            num_obs = len(y)
            y = np.asarray(y)[:, None]
        else:
            num_obs = 1
            y = np.array([y])[:, None]
        if num_obs == 1:
            x = np.reshape(x, (1, self.x_dim))
        x = torch.tensor(np.asarray(x), requires_grad=False).type(F64).to(self.device)
        y = torch.tensor(y, requires_grad=False).type(F64).to(self.device)
        if self.num_train == 0:
            self._X, self._y = x, y
        else:
            self._X = torch.cat((self._X, x), dim=0)
            self._y = torch.cat((self._y, y), dim=0)
        self.num_train += num_obs
        self.build_Ky_inv_mat()

    def build_Ky_inv_mat(self):
        """Gram matrix -> Cholesky -> Ky^-1, beta, moment-matching weights on the device
        (replaces the LU inverse of `src/gpr.py:159-171`)."""
        self._mats = {}
        lam, sf, nv = self._hyper_values()
        if self._owner is not None:
            self._owner._refit_member(self._index, lam, sf, nv)
            return
        if self._bundle is None:
            self._bundle = GPBundle(self.x_dim, 1, self.device.index or 0)
        self._bundle.fit(self._X, self._y, lam[None, :], [sf], [nv])

    # ---- materialised matrices ------------------------------------------------------------------
    def _matrix(self, which):
        if self.num_train == 0:
            return None
        if which not in self._mats:
            bundle, a = self._bundle_and_index()
            self._mats[which] = bundle.matrix(which, a)
        return self._mats[which]

    Kf = property(lambda self: self._matrix(_lib.MAT_KF))
    Ky = property(lambda self: self._matrix(_lib.MAT_KY))
    Ky_inv = property(lambda self: self._matrix(_lib.MAT_KY_INV))

    # ---- kernel / posterior -------------------------------------------------------------------
    def se_kernel(self, x1, x2):
        """k(x1, x2) for a single pair (`src/gpr.py:124-135`); tiny host-side torch expression."""
        lambdas = torch.exp(self.log_lambdas)
        d = torch.squeeze(x1) - torch.squeeze(x2)
        return torch.exp(self.log_sigma_f) ** 2 * torch.exp(-0.5 * torch.sum(d * d / lambdas))

    def compute_pred_train_covariance(self, X_pred):
        """K(X*, X_train): (p, n) tensor for 2-D input, (n,) for a single point (`src/gpr.py:253-283`)."""
        bundle, a = self._bundle_and_index()
        Xp = np.asarray(X_pred, dtype=np.float64)
        single = Xp.ndim == 1
        K = bundle.kernel_matrix(a, Xp.reshape(-1, self.x_dim), hyp=self._call_time_hypers())
        return K[0] if single else K

    def predict_latent_vars(self, X_pred, covar=False, targets=False):
        """Posterior mean (p,1) and covariance (p,p) as NumPy arrays (`src/gpr.py:285-332`)."""
        bundle, a = self._bundle_and_index()
        Xp = np.asarray(X_pred, dtype=np.float64).reshape(-1, self.x_dim)
        if self.f_nom is not None:
            # residual form K* Ky^-1 (y - f_nom(X)) + f_nom(X*)  (`src/gpr.py:309`): the nominal model is a Python
            # callable, so it is evaluated here; the products with K* and Ky^-1 run in the library
            Xt = torch.tensor(Xp, device=self.device).type(F64)
            resid = (self.y_train - self.f_nom(self.X_train)).detach().reshape(-1)
            mean, cov = bundle.predict(a, Xp, covar, targets, resid=resid, hyp=self._call_time_hypers())
            f_pred = torch.tensor(mean, device=self.device).type(F64)[:, None] + self.f_nom(Xt)
            return f_pred.cpu().detach().numpy(), (cov if covar else None)
        mean, cov = bundle.predict(a, Xp, covar, targets, hyp=self._call_time_hypers())
        return mean[:, None], (cov if covar else None)

    # ---- hyper-parameter training (src/gpr.py:240-251,334-370) ---------------------------------------
    def compute_marginal_likelihood(self):
        """Log marginal likelihood as a (1,1) tensor; differentiable w.r.t. the three log hyper-parameters."""
        if self.num_train == 0:
            raise RuntimeError("no training data")
        return _MarginalLikelihood.apply(self, self.log_lambdas, self.log_sigma_f, self.log_sigma_n)

    def update_hyperparams(self, num_iters=1000, verbose=True):
        """Adam ascent on the log marginal likelihood with the reference's schedule: step, rebuild the matrices,
        stop when every gradient component is below 1e-5."""
        for it in range(num_iters):
            self.optimizer.zero_grad()
            ml = self.compute_marginal_likelihood()
            ml.backward()
            self.optimizer.step()
            self.build_Ky_inv_mat()
            g_lam = self.log_lambdas.grad.cpu().detach().numpy()
            g_sf = self.log_sigma_f.grad.item()
            g_sn = self.log_sigma_n.grad.item()
            if verbose:
                print('Iter: ', it)
                print('ml: ', ml.item())
                print('lambdas: ', self.get_lambdas())
                print('sigma_f: ', self.get_sigma_f())
                print('sigma_n: ', self.get_sigma_n())
                print('log_lambdas.grad: ', g_lam)
                print('log_sigma_f.grad: ', g_sf)
                print('log_sigma_n.grad: ', g_sn)
                print('----------------------------------------')
            if (np.abs(g_lam) < 1e-5).all() and np.abs(g_sf) < 1e-5 and np.abs(g_sn) < 1e-5:
                break

    # ---- not provided -------------------------------------------------------------------------------
    def _not_provided(self, *_a, **_k):
        raise NotImplementedError("not provided: the reference marks this method as wrong / unused "
                                  "(src/gpr.py:139,204); the analytic gradient lives in libgpmpc "
                                  "(gpmpc_marginal_likelihood)")

    update_Ky_inv_mat = _not_provided
    kernel_matrix_gradient = _not_provided
    marginal_likelihood_grad = _not_provided
