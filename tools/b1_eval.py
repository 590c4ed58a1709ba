"""B=1 evaluation latency probe (development tool): python tools/b1_eval.py [reps]"""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpmpc_b200 as gp
n, H, E, m = 4096, 30, 4, 1
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
rng = np.random.default_rng(0)
S = rng.uniform(-1, 1, (n, E)); A = rng.uniform(-1, 1, (n, m))
nxt = 0.9 * S + 0.2 * np.tanh(np.concatenate([S, A], 1) @ rng.normal(0, 0.3, (E + m, E)))
dyn = gp.Dynamics(E, m)
for a in range(E):
    dyn.gpr_err[a].set_lambdas(np.full(E + m, 2.0)); dyn.gpr_err[a].set_sigma_n(np.float64(0.1))
dyn.append_train_data(S, A, nxt)
br = gp.BatchedRollouts(dyn, 2 * np.eye(E), 0.01 * np.eye(m))
U = rng.uniform(-0.3, 0.3, (1, H, m))
ts = []
for _ in range(reps):
    t0 = time.perf_counter()
    c, g = br.cost_and_grad(np.zeros(E), U, -1.0, host_out=True)
    ts.append(time.perf_counter() - t0)
print(c, "ms per eval:", [round(1e3 * t, 3) for t in ts[-5:]], "median", round(1e3 * float(np.median(ts)), 3))
