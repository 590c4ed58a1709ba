"""B=1 evaluation latency probe (development tool): python tools/b1_eval.py [reps] [n] [H]
Times one objective+gradient evaluation of a single control sequence (and of 2, 8, 32)."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpmpc_b200 as gp
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
H = int(sys.argv[3]) if len(sys.argv) > 3 else 30
E, m = 4, 1
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
print(f"n={n} H={H}: cost {c[0]:.12g}  ms per eval: last {[round(1e3 * t, 3) for t in ts[-3:]]} median {1e3 * float(np.median(ts)):.3f}")
for B in (2, 8, 32):
    Ub = rng.uniform(-0.3, 0.3, (B, H, m))
    ts = []
    for _ in range(max(3, reps // 4)):
        t0 = time.perf_counter(); br.cost_and_grad(np.zeros(E), Ub, -1.0, host_out=True); ts.append(time.perf_counter() - t0)
    print(f"  B={B}: {1e3 * float(np.median(ts)):.3f} ms per evaluation of the batch")
