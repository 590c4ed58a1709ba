"""Full-covariance rollout probe (development tool): python tools/fullcov_eval.py [n] [B] [H] [reps]
One cost+gradient evaluation of B control sequences under the full-covariance rollout (BASELINE config 4 shape)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gpmpc_b200 as gp
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
B = int(sys.argv[2]) if len(sys.argv) > 2 else 128
H = int(sys.argv[3]) if len(sys.argv) > 3 else 20
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 1
E, m = 4, 1
rng = np.random.default_rng(0)
S = rng.uniform(-1, 1, (n, E)); A = rng.uniform(-1, 1, (n, m))
nxt = 0.9 * S + 0.2 * np.tanh(np.concatenate([S, A], 1) @ rng.normal(0, 0.3, (E + m, E)))
dyn = gp.Dynamics(E, m)
for a in range(E):
    dyn.gpr_err[a].set_lambdas(np.full(E + m, 2.0)); dyn.gpr_err[a].set_sigma_n(np.float64(0.1))
dyn.append_train_data(S, A, nxt)
x0 = torch.tensor(rng.uniform(-0.5, 0.5, (B, E)), device="cuda"); U = torch.tensor(rng.uniform(-0.3, 0.3, (B, H, m)), device="cuda")
g = torch.full((B,), -1.0, dtype=torch.float64, device="cuda")
Q = 2 * np.eye(E); R = 0.01 * np.eye(m)
for full in (True, False):
    ts = []
    for _ in range(reps + 1):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        c, gr, _, _ = dyn._bundle.cost_grad(x0, U, g, Q, R, host_out=False, full=full)
        torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    print(f"n={n} B={B} H={H} full={full}: {min(ts[1:]) if reps else ts[0]:.3f} s per evaluation of the batch = "
          f"{B / (min(ts[1:]) if reps else ts[0]):.1f} evals/s; cost[0]={float(c[0]):.12g}")
