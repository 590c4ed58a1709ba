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
                 group=None):
        self.dynamics = dynamics
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
                                                 self.u_ref, want_grad=want_grad, host_out=host_out)
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
