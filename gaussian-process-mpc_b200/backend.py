"""GPBundle: thin Python owner of one libgpmpc handle (E GPs sharing the training inputs, one device).

PyTorch is used here only as plumbing: device buffers, the current CUDA stream and dtype conversion.
All arithmetic happens inside libgpmpc.so.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import _lib
from ._lib import GpmpcError, check

F64 = torch.float64


def require_cuda():
    if not torch.cuda.is_available():
        raise GpmpcError("a CUDA device is required: the GP-MPC hot path has no CPU fallback")


def _ptr(x):
    """Raw pointer of a contiguous fp64 torch tensor / numpy array (None -> NULL)."""
    if x is None:
        return None
    if isinstance(x, torch.Tensor):
        assert x.dtype == F64 and x.is_contiguous()
        return ctypes.c_void_p(x.data_ptr())
    assert isinstance(x, np.ndarray) and x.dtype == np.float64 and x.flags["C_CONTIGUOUS"]
    return ctypes.c_void_p(x.ctypes.data)


def as_f64(x, device=None):
    """Contiguous fp64 array: torch tensors stay where they are (or move to `device`), everything else
    becomes a C-contiguous numpy array (host pointer; the library stages it)."""
    if isinstance(x, torch.Tensor):
        t = x.detach().to(dtype=F64)
        if device is not None:
            t = t.to(device)
        return t.contiguous()
    return np.ascontiguousarray(np.asarray(x, dtype=np.float64))


class GPBundle:
    def __init__(self, D, E, device_index=0):
        require_cuda()
        self.lib = _lib.load()
        self.D, self.E, self.m = int(D), int(E), int(D) - int(E)
        self.device_index = int(device_index)
        self.device = torch.device("cuda", self.device_index)
        h = ctypes.c_void_p()
        rc = self.lib.gpmpc_create(self.device_index, self.D, self.E, ctypes.byref(h))
        if rc != 0:
            raise GpmpcError(f"gpmpc_create failed ({rc}): {self.lib.gpmpc_last_error(None).decode()}")
        self.h = h
        self.n = 0

    def __del__(self):
        h = getattr(self, "h", None)
        if h:
            try:
                self.lib.gpmpc_destroy(h)
            except Exception:
                pass
            self.h = None

    # ------------------------------------------------------------------------------------
    def _sync_stream(self):
        st = torch.cuda.current_stream(self.device).cuda_stream
        check(self.h, self.lib.gpmpc_set_stream(self.h, ctypes.c_void_p(st)), "gpmpc_set_stream")

    # ---- one rollout split over several GPUs (include/gpmpc.h: gpmpc_split_*) -------------------------------
    def split_connect(self, group=None):
        """One process per GPU (torch.distributed): exchange the mailboxes' CUDA IPC handles and connect.  Afterwards a
        B = 1 evaluation that EVERY rank calls with the same arguments is split over the ranks' GPUs (each must have
        fitted the same data); the kernels exchange their per-step sums over NVLink themselves."""
        import torch.distributed as dist
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        buf = (ctypes.c_ubyte * 64)()
        check(self.h, self.lib.gpmpc_split_export(self.h, ctypes.cast(buf, ctypes.c_void_p)), "gpmpc_split_export")
        on_gpu = dist.get_backend(group) == "nccl"
        mine = torch.tensor(list(buf), dtype=torch.uint8, device=self.device if on_gpu else "cpu")
        allh = torch.empty(world * 64, dtype=torch.uint8, device=mine.device)
        dist.all_gather_into_tensor(allh, mine, group=group)
        raw = bytes(allh.cpu().numpy().tobytes())
        dist.barrier(group)                      # every rank has created (and zeroed) its mailbox before anyone connects
        check(self.h, self.lib.gpmpc_split_connect(self.h, rank, world, ctypes.c_char_p(raw)), "gpmpc_split_connect")
        dist.barrier(group)

    @staticmethod
    def split_connect_local(bundles):
        """Bundles on different devices of THIS process (one host thread per bundle must then call concurrently)."""
        world = len(bundles)
        arr = (ctypes.c_void_p * world)(*[b.h for b in bundles])
        for r, b in enumerate(bundles):
            check(b.h, b.lib.gpmpc_split_connect_local(b.h, r, world, arr), "gpmpc_split_connect_local")

    def split_last_exchange_us(self):
        a = ctypes.c_double(0.0); b = ctypes.c_double(0.0)
        check(self.h, self.lib.gpmpc_split_last_exchange_us(self.h, ctypes.byref(a), ctypes.byref(b)), "split timeline")
        return a.value, b.value

    def split_disconnect(self):
        check(self.h, self.lib.gpmpc_split_disconnect(self.h), "gpmpc_split_disconnect")

    def set_option(self, name, value):
        check(self.h, self.lib.gpmpc_set_option(self.h, name.encode(), int(value)), "gpmpc_set_option")

    def synchronize(self):
        check(self.h, self.lib.gpmpc_synchronize(self.h), "gpmpc_synchronize")

    def fit(self, X, Y, lambdas, sigma_f, noise_var):
        self._sync_stream()
        X = as_f64(X); Y = as_f64(Y)
        n = X.shape[0]
        assert X.shape == (n, self.D) and Y.shape == (n, self.E)
        lam = as_f64(lambdas).reshape(self.E, self.D)
        sf = as_f64(sigma_f).reshape(self.E)
        nv = as_f64(noise_var).reshape(self.E)
        check(self.h, self.lib.gpmpc_fit(self.h, n, _ptr(X), _ptr(Y), _ptr(lam), _ptr(sf), _ptr(nv)), "gpmpc_fit")
        self.n = n

    def append_point(self, x, y):
        """Bordered update for one observation; returns False when a full fit is required instead."""
        self._sync_stream()
        x = as_f64(x).reshape(self.D); y = as_f64(y).reshape(self.E)
        rc = self.lib.gpmpc_append_point(self.h, _ptr(x), _ptr(y))
        if rc == 1:
            return False
        check(self.h, rc, "gpmpc_append_point")
        self.n += 1
        return True

    def refit_output(self, a, y, lambdas_a, sigma_f_a, noise_var_a):
        self._sync_stream()
        y = None if y is None else as_f64(y).reshape(-1)
        lam = as_f64(lambdas_a).reshape(self.D)
        check(self.h, self.lib.gpmpc_refit_output(self.h, int(a), _ptr(y), _ptr(lam), float(sigma_f_a),
                                                  float(noise_var_a)), "gpmpc_refit_output")

    def set_propagation_hypers(self, lambdas, sigma_f):
        self._sync_stream()
        lam = as_f64(lambdas).reshape(self.E, self.D)
        sf = as_f64(sigma_f).reshape(self.E)
        check(self.h, self.lib.gpmpc_set_propagation_hypers(self.h, _ptr(lam), _ptr(sf)),
              "gpmpc_set_propagation_hypers")

    def matrix(self, which, a):
        self._sync_stream()
        shape = (self.n,) if which == _lib.MAT_BETA else (self.n, self.n)
        out = torch.empty(shape, dtype=F64, device=self.device)
        check(self.h, self.lib.gpmpc_get_matrix(self.h, int(which), int(a), _ptr(out)), "gpmpc_get_matrix")
        return out

    def kernel_matrix(self, a, Xs, hyp=None):
        """K(X*, X) [p,n]; hyp = [lambdas (D), sigma_f, noise_var] overrides the fit-time kernel hyper-parameters."""
        self._sync_stream()
        Xs = as_f64(Xs)
        p = Xs.shape[0]
        out = torch.empty((p, self.n), dtype=F64, device=self.device)
        hp = None if hyp is None else as_f64(hyp).reshape(self.D + 2)
        check(self.h, self.lib.gpmpc_kernel_matrix_ex(self.h, int(a), p, _ptr(Xs), _ptr(hp), _ptr(out)),
              "gpmpc_kernel_matrix_ex")
        return out

    def predict(self, a, Xs, want_cov, add_noise, resid=None, hyp=None):
        """Posterior mean [p] (K* Ky^-1 resid if a residual vector y - f_nom(X) is given, else K* Ky^-1 y) and cov [p,p];
        hyp = [lambdas (D), sigma_f, noise_var] overrides the fit-time kernel hyper-parameters of K* and K**."""
        self._sync_stream()
        Xs = as_f64(Xs)
        p = Xs.shape[0]
        mean = np.empty(p)
        cov = np.empty((p, p)) if want_cov else None
        r = None if resid is None else as_f64(resid).reshape(-1)
        assert r is None or r.shape[0] == self.n
        hp = None if hyp is None else as_f64(hyp).reshape(self.D + 2)
        check(self.h, self.lib.gpmpc_predict_ex(self.h, int(a), p, _ptr(Xs), _ptr(r), _ptr(hp), _ptr(mean), _ptr(cov),
                                                int(bool(add_noise))), "gpmpc_predict_ex")
        return mean, cov

    def marginal_likelihood(self, a, resid=None, want_grad=True):
        """ml (float) and d ml / d [log lambdas, log sigma_f, log sigma_n] (numpy [D+2])."""
        self._sync_stream()
        r = None if resid is None else as_f64(resid).reshape(-1)
        ml = np.empty(1)
        grad = np.empty(self.D + 2) if want_grad else None
        check(self.h, self.lib.gpmpc_marginal_likelihood(self.h, int(a), _ptr(r), _ptr(ml), _ptr(grad)),
              "gpmpc_marginal_likelihood")
        return float(ml[0]), grad

    def moment_match(self, U, S, out_device=True):
        """U [B,D]; S [B,D] (diagonal) or [B,D,D] (full).  Returns mean [B,E], var [B,E]."""
        self._sync_stream()
        U = as_f64(U); S = as_f64(S)
        B = U.shape[0]
        full = S.ndim == 3
        if out_device:
            mean = torch.empty((B, self.E), dtype=F64, device=self.device)
            var = torch.empty((B, self.E), dtype=F64, device=self.device)
        else:
            mean = np.empty((B, self.E)); var = np.empty((B, self.E))
        check(self.h, self.lib.gpmpc_moment_match(self.h, B, _ptr(U), _ptr(S), int(full), _ptr(mean), _ptr(var)),
              "gpmpc_moment_match")
        return mean, var

    def moment_match_cov(self, U, S):
        """U [B,D], full S [B,D,D] -> mean [B,E], full output covariance [B,E,E] (numpy)."""
        self._sync_stream()
        U = as_f64(U); S = as_f64(S)
        B = U.shape[0]
        mean = np.empty((B, self.E)); cov = np.empty((B, self.E, self.E))
        check(self.h, self.lib.gpmpc_moment_match_cov(self.h, B, _ptr(U), _ptr(S), _ptr(mean), _ptr(cov)),
              "gpmpc_moment_match_cov")
        return mean, cov

    def rollout(self, x0, U, out_device=True):
        """x0 [B,E], U [B,H,m] -> means [B,H+1,E], vars [B,H+1,E]; keeps the tape for rollout_vjp."""
        self._sync_stream()
        x0 = as_f64(x0); U = as_f64(U)
        B, H = U.shape[0], U.shape[1]
        if out_device:
            means = torch.empty((B, H + 1, self.E), dtype=F64, device=self.device)
            vars_ = torch.empty((B, H + 1, self.E), dtype=F64, device=self.device)
        else:
            means = np.empty((B, H + 1, self.E)); vars_ = np.empty((B, H + 1, self.E))
        check(self.h, self.lib.gpmpc_rollout(self.h, B, H, _ptr(x0), _ptr(U), _ptr(means), _ptr(vars_)), "gpmpc_rollout")
        return means, vars_

    def rollout_vjp(self, B, H, gmeans, gvars, want_gx0=False):
        self._sync_stream()
        gm = None if gmeans is None else as_f64(gmeans)
        gv = None if gvars is None else as_f64(gvars)
        gU = torch.empty((B, H, self.m), dtype=F64, device=self.device)
        gx0 = torch.empty((B, self.E), dtype=F64, device=self.device) if want_gx0 else None
        check(self.h, self.lib.gpmpc_rollout_vjp(self.h, B, H, _ptr(gm), _ptr(gv), _ptr(gU), _ptr(gx0)),
              "gpmpc_rollout_vjp")
        return gU, gx0

    def rollout_full(self, x0, U, out_device=True):
        """Full-covariance rollout: x0 [B,E], U [B,H,m] -> means [B,H+1,E], covs [B,H+1,E,E]; keeps the tape for
        rollout_full_vjp."""
        self._sync_stream()
        x0 = as_f64(x0); U = as_f64(U)
        B, H = U.shape[0], U.shape[1]
        if out_device:
            means = torch.empty((B, H + 1, self.E), dtype=F64, device=self.device)
            covs = torch.empty((B, H + 1, self.E, self.E), dtype=F64, device=self.device)
        else:
            means = np.empty((B, H + 1, self.E)); covs = np.empty((B, H + 1, self.E, self.E))
        check(self.h, self.lib.gpmpc_rollout_full(self.h, B, H, _ptr(x0), _ptr(U), _ptr(means), _ptr(covs)), "gpmpc_rollout_full")
        return means, covs

    def rollout_full_vjp(self, B, H, gmeans, gcovs, want_gx0=False):
        self._sync_stream()
        gm = None if gmeans is None else as_f64(gmeans)
        gc = None if gcovs is None else as_f64(gcovs)
        gU = torch.empty((B, H, self.m), dtype=F64, device=self.device)
        gx0 = torch.empty((B, self.E), dtype=F64, device=self.device) if want_gx0 else None
        check(self.h, self.lib.gpmpc_rollout_full_vjp(self.h, B, H, _ptr(gm), _ptr(gc), _ptr(gU), _ptr(gx0)),
              "gpmpc_rollout_full_vjp")
        return gU, gx0

    def cost_grad(self, x0, U, gamma, Q, R, R_delta=None, last_u=None, x_ref=None, u_ref=None, want_grad=True,
                  want_traj=False, host_out=True, full=False):
        """Fused objective + gradient.  x0 [B,E], U [B,H,m], gamma [B].  Host (numpy) inputs are staged by
        the library; outputs are numpy when host_out else device tensors.  full=True propagates the full E x E state
        covariance (the 4th return value is then covs [B,H+1,E,E] instead of variances [B,H+1,E])."""
        self._sync_stream()
        x0 = as_f64(x0); U = as_f64(U); gamma = as_f64(gamma)
        B, H = U.shape[0], U.shape[1]
        Q = as_f64(Q).reshape(self.E, self.E)
        R = as_f64(R).reshape(self.m, self.m)
        Rd = None if R_delta is None else as_f64(R_delta).reshape(self.m, self.m)
        lu = None
        if Rd is not None:
            lu = as_f64(last_u).reshape(B, self.m)
        xr = np.zeros(self.E) if x_ref is None else as_f64(x_ref).reshape(self.E)
        ur = np.zeros(self.m) if u_ref is None else as_f64(u_ref).reshape(self.m)
        vshape = (B, H + 1, self.E, self.E) if full else (B, H + 1, self.E)
        if host_out:
            cost = np.empty(B)
            grad = np.empty((B, H, self.m)) if want_grad else None
            means = np.empty((B, H + 1, self.E)) if want_traj else None
            vars_ = np.empty(vshape) if want_traj else None
        else:
            cost = torch.empty(B, dtype=F64, device=self.device)
            grad = torch.empty((B, H, self.m), dtype=F64, device=self.device) if want_grad else None
            means = torch.empty((B, H + 1, self.E), dtype=F64, device=self.device) if want_traj else None
            vars_ = torch.empty(vshape, dtype=F64, device=self.device) if want_traj else None
        fn = self.lib.gpmpc_rollout_cost_grad_full if full else self.lib.gpmpc_rollout_cost_grad
        check(self.h, fn(self.h, B, H, _ptr(x0), _ptr(U), _ptr(gamma), _ptr(Q), _ptr(R), _ptr(Rd), _ptr(lu), _ptr(xr),
                         _ptr(ur), _ptr(cost), _ptr(grad), _ptr(means), _ptr(vars_)),
              "gpmpc_rollout_cost_grad_full" if full else "gpmpc_rollout_cost_grad")
        return cost, grad, means, vars_

    # ------------------------------------------------------------------------------------
    def launch_count(self):
        return int(self.lib.gpmpc_launch_count(self.h))

    def pair_kernel_timing(self):
        ms = ctypes.c_double(0.0); ev = ctypes.c_longlong(0)
        check(self.h, self.lib.gpmpc_last_pair_kernel_ms(self.h, ctypes.byref(ms), ctypes.byref(ev)), "timing")
        return ms.value, ev.value

    def set_pair_timing(self, on):
        """Armed timers synchronise the stream after every horizon step: disarm them after a measurement."""
        check(self.h, self.lib.gpmpc_set_pair_timing(self.h, int(bool(on))), "timing")

    def measure_fp64_peak(self):
        a = ctypes.c_double(0.0); b = ctypes.c_double(0.0)
        self._sync_stream()
        check(self.h, self.lib.gpmpc_measure_fp64_peak(self.h, ctypes.byref(a), ctypes.byref(b)), "peak")
        return a.value, b.value


_default = {}


def default_bundle(device_index=0):
    """A stateless handle (D=E=1) used by the free functions of uncertainty_prop."""
    b = _default.get(device_index)
    if b is None:
        b = GPBundle(1, 1, device_index)
        _default[device_index] = b
    return b
