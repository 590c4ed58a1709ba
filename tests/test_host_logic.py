"""CPU-only tests: the C-ABI library loads and exports every symbol include/gpmpc.h declares, the host-side
logic of the reference-shaped classes, and the multi-rank sharding/gather (gloo, world_size 2)."""
import os
import re
import socket
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as ge
    ge._load_build_module().build()
    from gpmpc_b200 import _lib
    header = open(os.path.join(ROOT, "include", "gpmpc.h")).read()
    declared = set(re.findall(r"\b(gpmpc_[a-z0-9_]+)\s*\(", header))
    declared.discard("gpmpc_ctx")
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.gpmpc_version() >= 100


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "gaussian-process-mpc_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "oracle" not in src.replace("test oracle", ""), fn


def test_no_cuda_means_loud_failure():
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    import gpmpc_b200 as gp
    dyn = gp.Dynamics(2, 1)
    with pytest.raises(gp.GpmpcError):
        dyn.append_train_data(np.zeros(2), np.zeros(1), np.zeros(2))


def test_setters_round_like_the_reference():
    """src/gpr.py:59,72,85: Python floats go through an fp32 tensor, float64 ndarrays do not."""
    import gpmpc_b200 as gp
    g = gp.GaussianProcessRegression(3)
    g.set_lambdas([0.7, 1.3, 2.0])
    expect = np.exp(np.log(np.array([0.7, 1.3, 2.0], dtype=np.float32)).astype(np.float64))
    assert np.allclose(g.get_lambdas(), expect, rtol=1e-7) and not np.array_equal(g.get_lambdas(), [0.7, 1.3, 2.0])
    g.set_lambdas(np.array([0.7, 1.3, 2.0]))
    assert np.allclose(g.get_lambdas(), [0.7, 1.3, 2.0], rtol=1e-15)
    from gpmpc_b200.gpr import noise_variance
    assert noise_variance(0.1) == float(np.float32(0.1 ** 2))
    k0 = g._hyper_key()
    g.set_sigma_f(2.0)
    assert g._hyper_key() != k0


def test_stack_observations_layouts():
    """src/dynamics.py:49-60 and src/test/test_dynamics.py:20-73: single and batched observations."""
    from gpmpc_b200 import Dynamics
    x, y = Dynamics._stack_observations(np.array([1., 2.]), np.array([3.]), np.array([4., 5.]), 2, 1)
    assert x.tolist() == [[1., 2., 3.]] and y.tolist() == [[4., 5.]]
    st = np.arange(6.).reshape(3, 2)
    x, y = Dynamics._stack_observations(st, np.array([7., 8., 9.]), st + 1, 2, 1)
    assert x.shape == (3, 3) and x[:, 2].tolist() == [7., 8., 9.] and y.shape == (3, 2)
    x, y = Dynamics._stack_observations(st, np.array([[7.], [8.], [9.]]), st + 1, 2, 1)
    assert x.shape == (3, 3)


class _FakeBundle:
    """Stands in for the device: cost = sum(U^2) + sum(x0), grad = 2U; counts calls."""
    def __init__(self):
        self.calls = 0

    def cost_grad(self, x0, U, gamma, Q, R, R_delta=None, last_u=None, x_ref=None, u_ref=None, want_grad=True,
                  want_traj=False, host_out=True, full=False):
        self.calls += 1
        U = np.asarray(U); x0 = np.asarray(x0)
        return (U ** 2).sum(axis=(1, 2)) + x0.sum(axis=1), 2 * U, None, None

    def set_propagation_hypers(self, *a):
        pass


def _fake_mpc(H=3, m=2, E=2):
    import gpmpc_b200 as gp
    mpc = gp.RiskSensitiveMPC(-1.0, H, E, m, 2 * np.eye(E), 0.1 * np.eye(m))
    mpc.dynamics._bundle = _FakeBundle()
    mpc.dynamics._X = torch.zeros((1, E + m))
    mpc.dynamics._prop_key = tuple(g._hyper_key() for g in mpc.dynamics.gpr_err)
    mpc.curr_state = torch.tensor([1.0, 2.0], dtype=torch.float64)
    return mpc


def test_objective_gradient_cache_semantics():
    """objective(x) then gradient(x) costs ONE device evaluation; gradient at a new x recomputes
    (superset of src/mpc.py:245-255, which ignores x)."""
    mpc = _fake_mpc()
    x = np.arange(6.) * 0.1
    c = mpc.objective(x)
    assert c == pytest.approx((x ** 2).sum() + 3.0)
    g = mpc.gradient(x)
    assert mpc.dynamics._bundle.calls == 1 and g.shape == (3, 2) and np.allclose(g.reshape(-1), 2 * x)
    g2 = mpc.gradient(x + 1.0)
    assert mpc.dynamics._bundle.calls == 2 and np.allclose(g2.reshape(-1), 2 * (x + 1.0))
    assert mpc.constraints(x) == 0 and mpc.jacobian(x).shape == x.shape


def test_get_optimal_trajectory_without_data_and_with_fallback_solver():
    import gpmpc_b200 as gp
    mpc = gp.RiskSensitiveMPC(-1.0, 3, 2, 2, 2 * np.eye(2), 0.1 * np.eye(2))
    assert np.array_equal(mpc.get_optimal_trajectory(np.zeros(2)), np.zeros((3, 2)))   # src/mpc.py:285-289
    mpc = _fake_mpc()
    mpc.dynamics.gpr_err[0].num_train = 1
    mpc.set_lb([-1, -1]); mpc.set_ub([1, 1])
    traj = mpc.get_optimal_trajectory(np.array([1.0, 2.0]))
    assert traj.shape == (3, 2) and np.allclose(traj, 0.0, atol=1e-6)
    assert np.array_equal(mpc.last_traj.reshape(3, 2), traj)


def test_shard_range_partitions():
    from gpmpc_b200 import shard_range
    for B in (1, 7, 8, 1024, 1030):
        for world in (1, 2, 3, 8):
            spans = [shard_range(B, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _rank_main(rank, world, port, B, out):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from gpmpc_b200 import BatchedRollouts
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    seen = []

    def fake_eval(x0, U, gamma, last_u=None, host_out=True, want_grad=True):
        seen.append(U.shape[0])
        return (U ** 2).sum(axis=(1, 2)) * gamma + x0.sum(axis=1), 2 * U * gamma[:, None, None]

    br = BatchedRollouts(evaluate_fn=fake_eval)
    rng = np.random.default_rng(0)
    U = rng.normal(size=(B, 4, 2)); x0 = rng.normal(size=3); gamma = rng.normal(size=B)
    cost, grad = br.cost_and_grad_sharded(x0, U, gamma)
    exp_c = (U ** 2).sum(axis=(1, 2)) * gamma + x0.sum()
    ok = np.allclose(cost.numpy(), exp_c) and np.allclose(grad.numpy(), 2 * U * gamma[:, None, None])
    out[rank] = (ok, seen[0] if seen else 0)
    dist.destroy_process_group()


@pytest.mark.parametrize("B", [9, 16])
def test_sharded_gather_world2_gloo(B):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    out = ctx.Manager().dict()
    port = _free_port()
    procs = [ctx.Process(target=_rank_main, args=(r, 2, port, B, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert out[0][0] and out[1][0]
    assert out[0][1] + out[1][1] == B        # shards cover the batch exactly once


def _solve_rank_main(rank, world, port, B, out):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from gpmpc_b200 import BatchedRollouts, BatchedSolver, shard_indices
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    H, m = 3, 2
    n = H * m
    rng = np.random.default_rng(0)
    A = rng.normal(size=(B, n, n)); Q = np.einsum("bij,bkj->bik", A, A) + 0.5 * np.eye(n)
    c = rng.normal(size=(B, n)) * 3
    seen = []

    def fake(x0, U, gamma, last_u=None, host_out=True, want_grad=True):
        ids = np.rint(x0[:, 0]).astype(int)            # the problem index travels in x0
        seen.extend(ids.tolist())
        u = U.reshape(len(ids), n)
        QX = np.einsum("bij,bj->bi", Q[ids], u)
        return 0.5 * np.einsum("bi,bi->b", u, QX) + np.einsum("bi,bi->b", c[ids], u), (QX + c[ids]).reshape(U.shape)

    x0 = np.repeat(np.arange(B, dtype=np.float64)[:, None], 2, axis=1)
    mk = lambda: BatchedSolver(BatchedRollouts(evaluate_fn=fake), H, m, lb=[-1, -1], ub=[1, 1], max_iter=200, gtol=1e-8)   # noqa: E731
    sharded = mk().solve_sharded(x0, np.full(B, -1.0))
    mine = set(seen)
    single = mk().solve(x0, np.full(B, -1.0))
    ok = (np.allclose(sharded["cost"], single["cost"], rtol=0, atol=1e-9) and np.allclose(sharded["U"], single["U"], atol=1e-5)
          and bool(sharded["converged"].all()) and mine == set(shard_indices(B, world, rank).tolist()))
    out[rank] = ok
    dist.destroy_process_group()


@pytest.mark.parametrize("B", [7, 10])
def test_sharded_solver_world2_gloo(B):
    """BatchedSolver.solve_sharded: each rank solves only its interleaved shard, one all-gather returns every solution."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    out = ctx.Manager().dict()
    port = _free_port()
    procs = [ctx.Process(target=_solve_rank_main, args=(r, 2, port, B, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    assert out[0] and out[1]


def test_batched_solver_matches_lbfgsb_on_box_constrained_quadratics():
    """N1 driver (lock-step projected L-BFGS over many problems) against scipy on a fake evaluator."""
    from scipy.optimize import minimize
    from gpmpc_b200 import BatchedRollouts, BatchedSolver
    rng = np.random.default_rng(0)
    B, H, m = 6, 4, 2
    n = H * m
    A = rng.normal(size=(B, n, n)); Q = np.einsum("bij,bkj->bik", A, A) + 0.5 * np.eye(n)
    c = rng.normal(size=(B, n)) * 3

    def fake(x0, U, gamma, last_u=None, host_out=True, want_grad=True):
        # the solver evaluates only the problems that still need a trial point: x0[:, 0] carries the problem id
        ids = np.asarray(x0)[:, 0].astype(int)
        X = U.reshape(U.shape[0], -1)
        QX = np.einsum("bij,bj->bi", Q[ids], X)
        f = 0.5 * np.einsum("bi,bi->b", X, QX) + np.einsum("bi,bi->b", c[ids], X)
        nan_region = (ids == 3) & (np.abs(X).max(axis=1) > 0.9)     # a NaN region must be treated as a rejected step
        f = np.where(nan_region, np.nan, f)
        return f, (QX + c[ids]).reshape(U.shape)

    sol = BatchedSolver(BatchedRollouts(evaluate_fn=fake), H, m, lb=[-1, -1], ub=[1, 1], max_iter=300, gtol=1e-8)
    r = sol.solve(np.repeat(np.arange(B, dtype=np.float64)[:, None], 3, axis=1), np.full(B, -1.0))
    for b in range(B):
        if b == 3:
            assert np.isfinite(r["cost"][b]) and np.abs(r["U"][b]).max() <= 0.9 + 1e-12
            continue
        ref = minimize(lambda x: (0.5 * x @ Q[b] @ x + c[b] @ x, Q[b] @ x + c[b]), np.zeros(n), jac=True, method="L-BFGS-B",
                       bounds=[(-1, 1)] * n, options={"gtol": 1e-10, "ftol": 1e-15})
        assert r["converged"][b] and abs(r["cost"][b] - ref.fun) < 1e-8 and np.max(np.abs(r["U"][b].ravel() - ref.x)) < 1e-4


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs next to ours) must print ONE JSON line with the
    contract's keys; run here on a tiny instance."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--n", "96", "--H", "2"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-500:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["metric"] == "gp_mpc_rollout_cost_grad_evals_per_sec" and d["value"] > 0
    # the unmodified reference when oracle/_ref is staged (build() stages it wherever /root/reference exists), else the port
    staged = os.path.exists(os.path.join(root, "oracle", "_ref", "src", "mpc.py"))
    assert d["cpu_baseline"]["kind"] == ("reference" if staged else "port") and d["cpu_baseline"]["cores"] >= 1
    if staged:
        lin = d["cpu_baseline"]["linearity"]
        assert lin["t_H1_s"] > 0 and lin["t_H2_s"] > 0 and lin["t_full_s"] > 0
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0


def test_batched_simulator_closed_loop_on_a_fake_plant():
    """BatchedSimulator (the many-instance counterpart of src/simulator.py:37-60): solve -> apply first action -> next
    state, warm-started; on a linear plant with a quadratic stage cost every instance must be driven towards 0."""
    from gpmpc_b200 import BatchedRollouts, BatchedSolver, BatchedSimulator
    H, m, E, B = 4, 1, 1, 5
    a_coef, b_coef = 0.9, 0.5

    def fake(x0, U, gamma, last_u=None, host_out=True, want_grad=True):
        # cost = sum_t x_t^2 + 0.01 u_t^2 with x_{t+1} = a x_t + b u_t (exact gradient by the adjoint recursion)
        Bn = U.shape[0]
        xs = np.zeros((Bn, H + 1)); xs[:, 0] = x0[:, 0]
        for t in range(H):
            xs[:, t + 1] = a_coef * xs[:, t] + b_coef * U[:, t, 0]
        cost = (xs[:, 1:] ** 2).sum(1) + 0.01 * (U[:, :, 0] ** 2).sum(1)
        lam = np.zeros(Bn); grad = np.zeros_like(U)
        for t in range(H - 1, -1, -1):
            lam = 2 * xs[:, t + 1] + a_coef * lam
            grad[:, t, 0] = b_coef * lam + 0.02 * U[:, t, 0]
        return cost, grad

    solver = BatchedSolver(BatchedRollouts(evaluate_fn=fake), H, m, lb=[-1.0], ub=[1.0], max_iter=100, gtol=1e-8)
    sim = BatchedSimulator(solver, lambda x, u: a_coef * x + b_coef * u, num_iters=12)
    x0 = np.linspace(-1.5, 1.5, B)[:, None]
    out = sim.run(x0, np.full(B, -1.0))
    assert out["states"].shape == (13, B, E) and out["actions"].shape == (12, B, m) and out["solves"] == 12 * B
    assert np.all(np.abs(out["states"][-1]) < 1e-2) and np.all(np.abs(out["actions"]) <= 1.0 + 1e-12)
    far = np.abs(x0[:, 0]) > 0.1
    assert np.all(np.abs(out["states"][3, far, 0]) < 0.5 * np.abs(x0[far, 0]))                    # the controller, not the plant's 0.9


def _claim_slices(arrivals, P):
    """Mirror of the slice claim in csrc/mm_step_single.cuh: `arrivals` lists the SM of every CTA in the order their
    atomics are served.  Returns the slice each CTA ends up with."""
    half = P // 2
    per_sm, nbig, nsmall, out = {}, 0, 0, []
    for sm in arrivals:
        first = per_sm.get(sm, 0) == 0
        per_sm[sm] = per_sm.get(sm, 0) + 1
        if first:
            k = nbig; nbig += 1
            if k < half:
                out.append(k)
            else:
                out.append(half + nsmall); nsmall += 1
        else:
            j = nsmall; nsmall += 1
            if j < half:
                out.append(half + j)
            else:
                out.append(nbig); nbig += 1
    return out


def _slice_range(s, P, total, tiles_big):
    half = P // 2
    sm = s if s < half else s - half
    base, cnt = (0, tiles_big) if s < half else (tiles_big, total - tiles_big)
    return base + cnt * sm // half, base + cnt * (sm + 1) // half


def test_single_rollout_slice_claim_is_always_a_bijection():
    """The CTAs of the single-rollout step kernel claim their tile slices (first arrival on an SM: a big one).  Whatever
    the arrival order and however unevenly the hardware spreads the CTAs over the SMs, every slice is taken exactly
    once, and the slices partition the tile list."""
    rng = np.random.default_rng(5)
    sms = 148
    P = 2 * sms
    cases = [np.repeat(np.arange(sms), 2)]                                  # two per SM, in order
    cases.append(np.concatenate([np.arange(sms), np.arange(sms)]))          # all first arrivals, then all second ones
    for _ in range(40):
        a = np.repeat(np.arange(sms), 2); rng.shuffle(a); cases.append(a)   # two per SM, any order
    for _ in range(40):
        cases.append(rng.integers(0, sms, P))                               # 0..k CTAs per SM
    cases.append(np.zeros(P, dtype=int))                                    # everything on one SM
    cases.append(np.arange(P) % (P // 4))                                   # four per SM on a quarter of the SMs
    for a in cases:
        got = _claim_slices(list(a), P)
        assert sorted(got) == list(range(P))
    a = np.repeat(np.arange(sms), 2)
    got = _claim_slices(list(a), P)
    assert all(s < sms for s in got[0::2]) and all(s >= sms for s in got[1::2])   # first arrival big, second small
    for total in (296, 297, 528, 8256, 131328):
        for share in (500, 660, 900):
            tiles_big = total * share // 1000
            edges = sorted(_slice_range(s, P, total, tiles_big) for s in range(P))
            assert edges[0][0] == 0 and edges[-1][1] == total
            assert all(edges[i][1] == edges[i + 1][0] for i in range(P - 1))
            assert all(lo <= hi for lo, hi in edges)
