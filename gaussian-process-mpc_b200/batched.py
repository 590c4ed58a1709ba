"""Batched evaluation of many independent rollouts (multi-start control sequences, gamma sweeps, initial
states) and their partition over the GPUs of one node.

The reference evaluates one control sequence per call (`src/mpc.py:202-255`).  The device kernels advance B
rollouts in lock step, so callers that have many independent problems hand them over together.

Multi-GPU: one process per GPU (torch.distributed, NCCL).  The GP training set is replicated (every rank
fits the same bundle -- deterministic); rollouts are split into contiguous shards; the only communication is
ONE all-gather of [cost | grad] per evaluation (a few KB: latency bound, NVLink bandwidth is irrelevant).
"""
from __future__ import annotations

import numpy as np
import torch

from .backend import F64


def shard_indices(B: int, world: int, rank: int):
    """Interleaved shard of B problems for `rank`: problems rank, rank + world, ...  Used for SOLVES, whose cost
    varies from problem to problem (a gamma sweep laid out gamma-major would otherwise give one rank all the hard
    ones); rollout evaluations cost the same and use the contiguous `shard_range`."""
    return np.arange(int(rank), int(B), int(world))


def shard_range(B: int, world: int, rank: int):
    """Contiguous shard [lo, hi) of B rollouts for `rank` of `world`; the first B % world ranks get one more."""
    base, extra = divmod(int(B), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


class BatchedRollouts:
    """Cost + gradient for B control sequences at once.

    evaluate_fn(x0[b,E], U[b,H,m], gamma[b], last_u[b,m] | None) -> (cost[b], grad[b,H,m]) can be injected
    (the CPU tests of the sharding logic do that); by default it is the fused device call of a fitted
    `Dynamics`."""

    def __init__(self, dynamics=None, Q=None, R=None, R_delta=None, x_ref=None, u_ref=None, evaluate_fn=None,
                 group=None, full=False):
        self.dynamics = dynamics
        self.full = bool(full)          # propagate the full E x E state covariance (BASELINE config 4)
        self.Q, self.R, self.R_delta = Q, R, R_delta
        self.x_ref, self.u_ref = x_ref, u_ref
        self.group = group
        self._evaluate = evaluate_fn if evaluate_fn is not None else self._device_evaluate

    # ---- single-device path -----------------------------------------------------------------------
    def _device_evaluate(self, x0, U, gamma, last_u=None, host_out=False, want_grad=True):
        dyn = self.dynamics
        dyn._require_data()
        dyn._sync_propagation_hypers()
        cost, grad, _, _ = dyn._bundle.cost_grad(x0, U, gamma, self.Q, self.R, self.R_delta, last_u, self.x_ref,
                                                 self.u_ref, want_grad=want_grad, host_out=host_out, full=self.full)
        dyn._tape_serial += 1
        return cost, grad

    @staticmethod
    def _broadcast_inputs(x0, U, gamma):
        B = U.shape[0]
        lib = torch if isinstance(U, torch.Tensor) else np
        if x0.ndim == 1:
            x0 = x0[None, :].expand(B, -1).contiguous() if lib is torch else np.broadcast_to(x0, (B, x0.shape[0])).copy()
        if np.isscalar(gamma) or getattr(gamma, "ndim", 1) == 0:
            gamma = np.full(B, float(gamma))
        return x0, U, gamma

    def cost_and_grad(self, x0, U, gamma, last_u=None, host_out=False):
        """x0 [E] or [B,E]; U [B,H,m]; gamma scalar or [B].  Returns cost [B], grad [B,H,m]."""
        x0, U, gamma = self._broadcast_inputs(x0, U, gamma)
        return self._evaluate(x0, U, gamma, last_u, host_out=host_out)

    # ---- sharded over the ranks of a process group -------------------------------------------------------
    def cost_and_grad_sharded(self, x0, U, gamma, last_u=None):
        """Every rank passes the SAME full batch; each evaluates its contiguous shard and one all-gather
        returns the full cost [B] and grad [B,H,m] on every rank (device tensors with NCCL, CPU with gloo)."""
        import torch.distributed as dist
        x0, U, gamma = self._broadcast_inputs(x0, U, gamma)
        world = dist.get_world_size(self.group)
        rank = dist.get_rank(self.group)
        B, H, m = U.shape
        lo, hi = shard_range(B, world, rank)
        per = (B + world - 1) // world
        lu = None if last_u is None else last_u[lo:hi]
        on_gpu = dist.get_backend(self.group) == "nccl"
        dev = torch.device("cuda", torch.cuda.current_device()) if on_gpu else torch.device("cpu")
        packed = torch.zeros((per, 1 + H * m), dtype=F64, device=dev)
        if hi > lo:
            cost, grad = self._evaluate(x0[lo:hi], U[lo:hi], gamma[lo:hi], lu, host_out=not on_gpu)
            cost = torch.as_tensor(cost, dtype=F64, device=dev)
            grad = torch.as_tensor(grad, dtype=F64, device=dev)
            packed[:hi - lo, 0] = cost
            packed[:hi - lo, 1:] = grad.reshape(hi - lo, H * m)
        gathered = torch.empty((world, per, 1 + H * m), dtype=F64, device=dev)
        dist.all_gather_into_tensor(gathered.view(world * per, 1 + H * m), packed, group=self.group)
        rows = []
        for r in range(world):
            rlo, rhi = shard_range(B, world, r)
            rows.append(gathered[r, :rhi - rlo])
        full = torch.cat(rows, dim=0)
        return full[:, 0].contiguous(), full[:, 1:].reshape(B, H, m).contiguous()


class BatchedSolver:
    """Lock-step solver for MANY independent MPC problems (multi-start control sequences, gamma sweeps, many
    initial states): SURVEY 8f row N1.  The reference solves one NLP at a time through cyipopt
    (`src/mpc.py:269-330`), which keeps a GPU at B = 1; here every iteration is ONE batched device evaluation
    of cost and gradient for all still-active problems.

    Method: projected L-BFGS with Armijo backtracking on the box lb <= u <= ub (per-problem histories, all
    vectorised over the batch).  Non-finite costs (the NaN of log(det) < 0) are treated like IPOPT
    treats evaluation errors: the trial step is rejected and shortened.
    """

    def __init__(self, rollouts: BatchedRollouts, horizon, input_dim, lb=None, ub=None, memory=8, max_iter=100,
                 gtol=1e-4, ftol=1e-10, max_backtracks=12):
        self.rollouts = rollouts
        self.H, self.m = int(horizon), int(input_dim)
        n = self.H * self.m
        self.lb = np.full(n, -np.inf) if lb is None else np.tile(np.asarray(lb, dtype=np.float64), self.H)
        self.ub = np.full(n, np.inf) if ub is None else np.tile(np.asarray(ub, dtype=np.float64), self.H)
        self.memory, self.max_iter, self.gtol, self.ftol = memory, max_iter, gtol, ftol
        self.max_backtracks = max_backtracks
        self.n_evals = 0              # batched device evaluations
        self.n_rollout_evals = 0      # rollouts evaluated in total (sum of the batch sizes)

    def _eval(self, x0, X, gamma, last_u):
        B = X.shape[0]
        cost, grad = self.rollouts.cost_and_grad(x0, X.reshape(B, self.H, self.m), gamma, last_u, host_out=True)
        self.n_evals += 1
        self.n_rollout_evals += B
        return np.asarray(cost, dtype=np.float64), np.asarray(grad, dtype=np.float64).reshape(B, -1)

    def _direction(self, G, S, Y, rho, count):
        """Two-loop recursion for every problem at once.  S, Y: [B, mem, n]; count: history length per problem."""
        B, mem, _ = S.shape
        q = G.copy()
        alpha = np.zeros((B, mem))
        for i in range(mem - 1, -1, -1):
            use = (i < count)
            a = rho[:, i] * np.einsum("bn,bn->b", S[:, i], q) * use
            alpha[:, i] = a
            q -= a[:, None] * Y[:, i]
        last = np.clip(count - 1, 0, mem - 1)
        idx = np.arange(B)
        yy = np.einsum("bn,bn->b", Y[idx, last], Y[idx, last])
        sy = np.einsum("bn,bn->b", S[idx, last], Y[idx, last])
        scale = np.where((count > 0) & (yy > 0), sy / np.where(yy > 0, yy, 1.0), 1.0)
        r = q * scale[:, None]
        for i in range(mem):
            use = (i < count)
            beta = rho[:, i] * np.einsum("bn,bn->b", Y[:, i], r) * use
            r += (alpha[:, i] - beta)[:, None] * S[:, i]
        return -r

    def solve(self, x0, gamma, U0=None, last_u=None):
        """x0 [B,E] (or [E]), gamma [B] (or scalar).  Returns dict(U [B,H,m], cost [B], iters, evals, converged [B])."""
        x0 = np.asarray(x0, dtype=np.float64)
        if x0.ndim == 1:
            B = int(np.size(gamma)) if np.ndim(gamma) else (1 if U0 is None else U0.shape[0])
            x0 = np.broadcast_to(x0, (B, x0.shape[0])).copy()
        B = x0.shape[0]
        gamma = np.broadcast_to(np.asarray(gamma, dtype=np.float64), (B,)).copy()
        n = self.H * self.m
        X = np.zeros((B, n)) if U0 is None else np.asarray(U0, dtype=np.float64).reshape(B, n).copy()
        X = np.clip(X, self.lb, self.ub)
        F, G = self._eval(x0, X, gamma, last_u)
        mem = self.memory
        S = np.zeros((B, mem, n)); Y = np.zeros((B, mem, n)); rho = np.zeros((B, mem))
        count = np.zeros(B, dtype=np.int64)
        active = np.isfinite(F)
        converged = np.zeros(B, dtype=bool)
        it = 0
        for it in range(1, self.max_iter + 1):
            # projected-gradient optimality measure
            pg = X - np.clip(X - G, self.lb, self.ub)
            done = np.max(np.abs(pg), axis=1) < self.gtol
            converged |= done & active
            active &= ~done
            if not active.any():
                break
            # active set: variables sitting on a bound with the gradient pushing outwards are frozen
            at_bound = ((X <= self.lb) & (G > 0)) | ((X >= self.ub) & (G < 0))
            Gf = np.where(at_bound, 0.0, G)
            Dir = self._direction(Gf, S, Y, rho, count)
            Dir[at_bound] = 0.0
            # fall back to (projected) steepest descent where the quasi-Newton direction is not a descent direction
            slope = np.einsum("bn,bn->b", Dir, Gf)
            bad = ~(slope < 0)
            Dir[bad] = -Gf[bad]
            # first trial: the unit quasi-Newton step; without curvature history (first iteration, or after a reset)
            # a steepest-descent step of unit length, like L-BFGS-B's first line search
            dnorm = np.linalg.norm(Dir, axis=1)
            step = np.where((count == 0) & (dnorm > 1.0), 1.0 / np.where(dnorm > 0, dnorm, 1.0), 1.0)
            Xn, Fn, Gn = X.copy(), F.copy(), G.copy()
            pending = active.copy()
            retried = np.zeros(B, dtype=bool)
            for _ in range(2 * self.max_backtracks):
                # only the problems that still need a trial point are evaluated (converged problems and accepted
                # steps cost nothing): the device batch shrinks as the line searches finish
                idx = np.where(pending)[0]
                T = np.clip(X[idx] + step[idx, None] * Dir[idx], self.lb, self.ub)
                Ft, Gt = self._eval(x0[idx], T, gamma[idx], None if last_u is None else np.asarray(last_u)[idx])
                dec = np.einsum("bn,bn->b", G[idx], T - X[idx])
                good = np.isfinite(Ft) & (Ft <= F[idx] + 1e-4 * dec) & (dec < 0)
                acc = idx[good]
                Xn[acc], Fn[acc], Gn[acc] = T[good], Ft[good], Gt[good]
                ok = np.zeros(B, dtype=bool); ok[acc] = True
                pending &= ~ok
                if not pending.any():
                    break
                step[pending] *= 0.5
                # a quasi-Newton direction that keeps failing is replaced once by the projected gradient
                retry = pending & (step < 0.5 ** self.max_backtracks) & ~retried
                if retry.any():
                    Dir[retry] = -Gf[retry]
                    step[retry] = 1.0
                    count[retry] = 0
                    retried |= retry
            stalled = pending                     # no acceptable step found: stop these problems
            s = Xn - X; y = Gn - G
            sy = np.einsum("bn,bn->b", s, y)
            upd = active & ~stalled & (sy > 1e-12)
            if upd.any():
                slot = np.minimum(count, mem - 1)
                full = upd & (count >= mem)
                if full.any():                    # drop the oldest pair
                    S[full, :-1] = S[full, 1:]; Y[full, :-1] = Y[full, 1:]; rho[full, :-1] = rho[full, 1:]
                idx = np.where(upd)[0]
                S[idx, slot[idx]] = s[idx]; Y[idx, slot[idx]] = y[idx]; rho[idx, slot[idx]] = 1.0 / sy[idx]
                count[idx] = np.minimum(count[idx] + 1, mem)
            small = np.abs(F - Fn) <= self.ftol * np.maximum(1.0, np.abs(F))
            X, F, G = Xn, Fn, Gn
            converged |= active & small & ~stalled
            active &= ~(stalled | small)
        return {"U": X.reshape(B, self.H, self.m), "cost": F, "iters": it, "evals": self.n_evals,
                "rollout_evals": self.n_rollout_evals, "converged": converged}

    def solve_sharded(self, x0, gamma, U0=None, last_u=None, group=None):
        """Many independent MPC problems over the ranks of a process group (BASELINE configs[4]: a gamma sweep x initial
        states "partitioned across the 8 GPUs"): every rank passes the SAME problem list, solves its interleaved shard
        with its own replica of the GP (no communication while solving: the problems are independent, so the lock
        step is per shard and a slow problem never stalls another GPU), and ONE all-gather returns U, cost and the
        convergence flags of all problems on every rank."""
        import torch.distributed as dist
        x0 = np.asarray(x0, dtype=np.float64)
        gamma = np.asarray(gamma, dtype=np.float64)
        B = x0.shape[0] if x0.ndim == 2 else int(gamma.size)
        if x0.ndim == 1:
            x0 = np.broadcast_to(x0, (B, x0.shape[0])).copy()
        gamma = np.broadcast_to(gamma, (B,)).copy()
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        idx = shard_indices(B, world, rank)
        n = self.H * self.m
        per = (B + world - 1) // world
        on_gpu = dist.get_backend(group) == "nccl"
        dev = torch.device("cuda", torch.cuda.current_device()) if on_gpu else torch.device("cpu")
        packed = torch.zeros((per, n + 4), dtype=F64)          # [U | cost | converged | iterations | rollout evals]
        if idx.size:
            sol = self.solve(x0[idx], gamma[idx], None if U0 is None else np.asarray(U0)[idx],
                             None if last_u is None else np.asarray(last_u)[idx])
            packed[:idx.size, :n] = torch.from_numpy(sol["U"].reshape(idx.size, n))
            packed[:idx.size, n] = torch.from_numpy(sol["cost"])
            packed[:idx.size, n + 1] = torch.from_numpy(sol["converged"].astype(np.float64))
            packed[:idx.size, n + 2] = float(sol["iters"])
            packed[:idx.size, n + 3] = float(sol["rollout_evals"]) / idx.size
        packed = packed.to(dev)
        gathered = torch.empty((world * per, n + 4), dtype=F64, device=dev)
        dist.all_gather_into_tensor(gathered, packed, group=group)
        g = gathered.cpu().numpy().reshape(world, per, n + 4)
        U = np.zeros((B, n)); cost = np.zeros(B); conv = np.zeros(B, dtype=bool); iters = np.zeros(world); revals = 0.0
        for r in range(world):
            ir = shard_indices(B, world, r)
            U[ir] = g[r, :ir.size, :n]; cost[ir] = g[r, :ir.size, n]; conv[ir] = g[r, :ir.size, n + 1] > 0.5
            if ir.size:
                iters[r] = g[r, 0, n + 2]; revals += g[r, 0, n + 3] * ir.size
        return {"U": U.reshape(B, self.H, self.m), "cost": cost, "converged": conv, "iters_per_rank": iters,
                "rollout_evals": revals}


class BatchedSimulator:
    """Closed loop for MANY independent MPC instances at once: the batched counterpart of the reference's
    `Simulator.run` (`src/simulator.py:37-60`: reset -> loop { solve, env.step(first action), append data }).

    Every iteration solves all instances in lock step with `BatchedSolver` (warm-started from the previous solution
    shifted by one step), applies the first action of each plan to the plant `step_fn(x[B,E], u[B,m]) -> x_next[B,E]`
    (a vectorised plant model; the reference's gym environments are out of scope) and, if `learn_every` > 0, appends
    the observed transitions of instance 0..k to the SHARED dynamics model every `learn_every` iterations (the
    reference appends after every step of its single instance; with many instances sharing one GP the model update
    is a batch append + refit).  Instances are independent, so under `torchrun` each rank runs its own shard."""

    def __init__(self, solver, step_fn, num_iters=50, learn_every=0, learn_instances=1):
        self.solver, self.step_fn = solver, step_fn
        self.num_iters, self.learn_every, self.learn_instances = int(num_iters), int(learn_every), int(learn_instances)

    def run(self, x0, gamma):
        """x0 [B,E], gamma [B] -> dict(states [T+1,B,E], actions [T,B,m], costs [T,B], solves, rollout_evals)."""
        x = np.array(x0, dtype=np.float64)
        B = x.shape[0]
        gamma = np.broadcast_to(np.asarray(gamma, dtype=np.float64), (B,)).copy()
        H, m = self.solver.H, self.solver.m
        states, actions, costs = [x.copy()], [], []
        U = np.zeros((B, H, m))
        buf_s, buf_a, buf_n = [], [], []
        e0 = self.solver.n_rollout_evals
        for it in range(self.num_iters):
            sol = self.solver.solve(x, gamma, U0=U)
            U = sol["U"]
            a = U[:, 0, :]
            xn = np.asarray(self.step_fn(x, a), dtype=np.float64).reshape(B, -1)
            actions.append(a.copy()); costs.append(sol["cost"].copy())
            if self.learn_every > 0:
                k = min(self.learn_instances, B)
                buf_s.append(x[:k].copy()); buf_a.append(a[:k].copy()); buf_n.append(xn[:k].copy())
                if (it + 1) % self.learn_every == 0:
                    dyn = self.solver.rollouts.dynamics
                    dyn.append_train_data(np.concatenate(buf_s), np.concatenate(buf_a), np.concatenate(buf_n))
                    buf_s, buf_a, buf_n = [], [], []
            x = xn
            states.append(x.copy())
            U = np.concatenate([U[:, 1:], U[:, -1:]], axis=1)          # shifted warm start
        return {"states": np.stack(states), "actions": np.stack(actions), "costs": np.stack(costs),
                "solves": B * self.num_iters, "rollout_evals": self.solver.n_rollout_evals - e0}
