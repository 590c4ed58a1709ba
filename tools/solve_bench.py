"""Latency of single-sequence evaluations (the cyipopt callback pattern) and of one NLP solve.
    python tools/solve_bench.py [n] [H]
"""
import sys, time, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gpmpc_b200 as gp

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
H = int(sys.argv[2]) if len(sys.argv) > 2 else 30
E, m = 4, 1
rng = np.random.default_rng(0)
S = rng.uniform(-1, 1, (n, E)); A = rng.uniform(-1, 1, (n, m))
W = rng.normal(0, 0.3, (E + m, E))
nxt = 0.9 * S + 0.2 * np.tanh(np.concatenate([S, A], 1) @ W)
mpc = gp.RiskSensitiveMPC(-1.0, H, E, m, 2 * np.eye(E), 0.01 * np.eye(m))
for a in range(E):
    mpc.dynamics.gpr_err[a].set_lambdas(np.full(E + m, 2.0)); mpc.dynamics.gpr_err[a].set_sigma_n(np.float64(0.1))
mpc.dynamics.append_train_data(S, A, nxt)
mpc.set_lb([-1.0]); mpc.set_ub([1.0])
mpc.curr_state = torch.tensor(rng.uniform(-0.5, 0.5, E), device="cuda")
x = rng.uniform(-0.3, 0.3, H * m)
for _ in range(3):
    mpc.objective(x); mpc.gradient(x)
ts = []
for i in range(20):
    xi = x + 1e-3 * i
    t0 = time.perf_counter(); c = mpc.objective(xi); g = mpc.gradient(xi); ts.append(time.perf_counter() - t0)
print(f"n={n} H={H}: objective+gradient (B=1) median {1e3*np.median(ts):.3f} ms, min {1e3*min(ts):.3f} ms, cost={c:.6f}")
st = []
for i in range(5):
    x0 = rng.uniform(-0.5, 0.5, E)
    n0 = mpc.n_evals
    t0 = time.perf_counter(); traj = mpc.get_optimal_trajectory(x0); st.append((time.perf_counter() - t0, mpc.n_evals - n0))
print("solve (L-BFGS-B fallback unless cyipopt present): " + ", ".join(f"{1e3*t:.1f} ms/{k} evals" for t, k in st),
      f"p50 {1e3*np.median([t for t, _ in st]):.1f} ms")
br = gp.BatchedRollouts(mpc.dynamics, 2 * np.eye(E), 0.01 * np.eye(m))
for B in (1, 4, 16, 48, 64, 128):
    U = rng.uniform(-0.3, 0.3, (B, H, m))
    for _ in range(2): br.cost_and_grad(np.zeros(E), U, -1.0, host_out=True)
    t0 = time.perf_counter(); br.cost_and_grad(np.zeros(E), U, -1.0, host_out=True); dt = time.perf_counter() - t0
    print(f"B={B}: {1e3*dt:.3f} ms  ({1e3*dt/B:.3f} ms per rollout)")
