"""Timings of the other BASELINE.json configurations (2, 4, 5) on one B200 (config 3 is bench.py).
    python tools/config_bench.py [2] [4] [5]
"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gpmpc_b200 as gp

which = set(sys.argv[1:]) or {"2", "4", "5"}


def synth(n, E, m, seed=0):
    rng = np.random.default_rng(seed)
    S = rng.uniform(-1, 1, (n, E)); A = rng.uniform(-1, 1, (n, m))
    nxt = 0.9 * S + 0.2 * np.tanh(np.concatenate([S, A], 1) @ rng.normal(0, 0.3, (E + m, E)))
    return S, A, nxt, rng


def dynamics(n, E, m, ard=False):
    S, A, nxt, rng = synth(n, E, m)
    dyn = gp.Dynamics(E, m)
    for a in range(E):
        dyn.gpr_err[a].set_lambdas(np.full(E + m, 2.0 + (0.1 * a if ard else 0.0))); dyn.gpr_err[a].set_sigma_n(np.float64(0.1))
    t0 = time.perf_counter(); dyn.append_train_data(S, A, nxt); dyn._bundle.synchronize()
    return dyn, rng, time.perf_counter() - t0


def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    return min(ts)


if "2" in which:      # batched moment matching: n=2048, D=5, E=4, 8192 uncertain inputs (mean + variance)
    dyn, rng, tf = dynamics(2048, 4, 1)
    U = torch.tensor(rng.uniform(-0.5, 0.5, (8192, 5)), device="cuda"); Sd = torch.tensor(rng.uniform(1e-3, 5e-2, (8192, 5)), device="cuda")
    t = timed(lambda: dyn._bundle.moment_match(U, Sd, out_device=True))
    pairs = 8192 * 4 * 2048 * 2049 / 2
    print(f"config 2: fit {tf:.3f} s; 8192 moment-matched predictions (mean+var, 4 outputs) in {1e3*t:.1f} ms = "
          f"{8192/t:.0f} inputs/s, {pairs/t/1e9:.0f} G (input, output, pair) evaluations/s")

if "4" in which:      # large exact GP n=16384, H=20
    dyn, rng, tf = dynamics(16384, 4, 1)
    br = gp.BatchedRollouts(dyn, 2 * np.eye(4), 0.01 * np.eye(1))
    B, H = 256, 20
    U = torch.tensor(rng.uniform(-0.3, 0.3, (B, H, 1)), device="cuda"); x0 = torch.tensor(rng.uniform(-0.5, 0.5, (B, 4)), device="cuda")
    g = torch.full((B,), -1.0, dtype=torch.float64, device="cuda")
    t = timed(lambda: dyn._bundle.cost_grad(x0, U, g, 2 * np.eye(4), 0.01 * np.eye(1), host_out=False), reps=1)
    print(f"config 4: fit n=16384 (4 outputs, shared hyper-parameters) {tf:.3f} s; variance-only rollout cost+gradient B={B} H={H}: "
          f"{t:.2f} s = {B/t:.1f} evals/s")
    t0 = time.perf_counter(); dyn.forward_propagate_full(2, rng.uniform(-0.5, 0.5, 4), rng.uniform(-0.3, 0.3, (2, 1))); t1 = time.perf_counter() - t0
    print(f"          full cross-output covariance rollout (forward values, generic kernels): {t1/2:.2f} s per horizon step")
    del dyn, br
    torch.cuda.empty_cache()

if "5" in which:      # gamma sweep x 512 initial states as lock-step MPC solves: n=4096, H=30
    dyn, rng, tf = dynamics(4096, 4, 1)
    br = gp.BatchedRollouts(dyn, 2 * np.eye(4), 0.01 * np.eye(1))
    gammas = np.array([-2.0, -1.0, 0.5, 1.0]); starts = rng.uniform(-0.5, 0.5, (128, 4))
    G, I = np.meshgrid(gammas, np.arange(128), indexing="ij")
    solver = gp.BatchedSolver(br, 30, 1, lb=[-1.0], ub=[1.0], max_iter=40, gtol=1e-4)
    t0 = time.perf_counter(); sol = solver.solve(starts[I.reshape(-1)], G.reshape(-1)); t = time.perf_counter() - t0
    print(f"config 5: {G.size} MPC instances (4 gammas x 128 initial states) solved in lock step: {t:.1f} s, {sol['iters']} iterations, "
          f"{sol['evals']} batched evaluations ({sol['rollout_evals'] / G.size:.1f} rollout evaluations per instance), "
          f"{int(sol['converged'].sum())} converged to gtol 1e-4 -> {G.size/t:.1f} solves/s")

if "n" in which:      # the "next" rows: incremental refit (N2) and hyper-parameter training steps (N3), n=4096
    dyn, rng, tf = dynamics(4096, 4, 1, ard=True)
    t0 = time.perf_counter(); dyn._fit_all(); dyn._bundle.synchronize(); t_full = time.perf_counter() - t0
    ts = []
    for _ in range(5):
        s = rng.uniform(-1, 1, 4); a = rng.uniform(-1, 1, 1)
        t0 = time.perf_counter(); dyn.append_train_data(s, a, 0.9 * s); dyn._bundle.synchronize(); ts.append(time.perf_counter() - t0)
    print(f"N2: full refit of 4 outputs (distinct hyper-parameters) at n=4096: {1e3*t_full:.1f} ms; one appended observation "
          f"(bordered update of Ky^-1, beta, Wt for 4 outputs): {1e3*np.median(ts):.2f} ms")
    g = dyn.gpr_err[0]
    t0 = time.perf_counter(); g.update_hyperparams(num_iters=10, verbose=False); dyn._bundle.synchronize(); t = time.perf_counter() - t0
    print(f"N3: 10 Adam steps on the log marginal likelihood of one output at n={g.num_train} (likelihood + analytic gradient + "
          f"refit per step): {1e3*t/10:.1f} ms per step")

if "1" in which:      # the reference's own experiment (src/experiments/pretrain_uncertainty.py): n=400, E=2, m=2, H=6, gamma=-1
    g = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "shipped.npz"))
    mpc = gp.RiskSensitiveMPC(-1, 6, 2, 2, 2 * np.identity(2), np.zeros((2, 2)), None)
    for i in range(2):
        mpc.dynamics.gpr_err[i].set_sigma_n(np.float64(g["ship_sn"][i]))
        mpc.dynamics.gpr_err[i].set_lambdas(np.asarray(g["ship_lam"][i], dtype=np.float64))
        mpc.dynamics.gpr_err[i].set_sigma_f(np.float64(g["ship_sf"][i]))
    t0 = time.perf_counter(); mpc.dynamics.append_train_data(g["ship_S"], g["ship_A"], g["ship_next"]); mpc.dynamics._bundle.synchronize()
    tf = time.perf_counter() - t0
    mpc.set_xref(np.array([0., 0.])); mpc.set_uref(np.array([0., 0.])); mpc.set_lb([-1.0, -1.0]); mpc.set_ub([1.0, 1.0])
    mpc.curr_state = torch.tensor(g["ship_x0"], device="cuda")
    x = g["ship_U0"].reshape(-1).copy()
    for _ in range(3): mpc.objective(x); mpc.gradient(x)
    ts = []
    for i in range(50):
        xi = x + 1e-4 * i
        t0 = time.perf_counter(); c = mpc.objective(xi); gr = mpc.gradient(xi); ts.append(time.perf_counter() - t0)
    st = []
    for i in range(5):
        n0 = mpc.n_evals; t0 = time.perf_counter(); mpc.get_optimal_trajectory(g["ship_x0"]); st.append((time.perf_counter() - t0, mpc.n_evals - n0))
    print(f"config 1 (shipped data, n=400, H=6): fit {1e3*tf:.1f} ms (first call), objective+gradient {1e3*np.median(ts):.3f} ms "
          f"(reference on CPU: 150 + 30 ms, SURVEY appendix C); one solve p50 {1e3*np.median([t for t, _ in st]):.1f} ms "
          f"({[k for _, k in st]} evaluations, L-BFGS-B fallback)")
