"""Fit + rollout at large n (config 4 scale): timing and identity residuals."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gpmpc_b200 as gp
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
ard = len(sys.argv) > 2 and sys.argv[2] == "ard"      # distinct hyper-parameters per output: no shared factorisation
E, m = 4, 1
rng = np.random.default_rng(0)
S = rng.uniform(-1, 1, (n, E)); A = rng.uniform(-1, 1, (n, m))
nxt = 0.9 * S + 0.2 * np.tanh(np.concatenate([S, A], 1) @ rng.normal(0, 0.3, (E + m, E)))
dyn = gp.Dynamics(E, m)
for a in range(E):
    dyn.gpr_err[a].set_lambdas(np.full(E + m, 2.0 + (0.1 * a if ard else 0.0))); dyn.gpr_err[a].set_sigma_n(np.float64(0.1))
t0 = time.perf_counter(); dyn.append_train_data(S, A, nxt); dyn._bundle.synchronize(); t1 = time.perf_counter()
print(f"n={n}: fit of {E} outputs {t1 - t0:.3f} s; mem {torch.cuda.memory_allocated()/1e9:.1f} GB torch + lib buffers")
t0 = time.perf_counter(); dyn.append_train_data(S, A, nxt); dyn._bundle.synchronize(); t1 = time.perf_counter()
print(f"second fit (2n points would be too big; refit same data again): n={dyn.gpr_err[0].num_train}")
g = dyn.gpr_err[1]
Kinv = g.Ky_inv; Ky = g.Ky
rows = torch.tensor(rng.integers(0, Kinv.shape[0], 8), device="cuda")
R = Kinv[rows] @ Ky
I = torch.zeros_like(R); I[torch.arange(8), rows] = 1.0
print("max |Kinv Ky - I| on 8 rows:", (R - I).abs().max().item())
del Kinv, Ky, R, I; g._mats = {}
torch.cuda.empty_cache()
br = gp.BatchedRollouts(dyn, 2 * np.eye(E), 0.01 * np.eye(m))
B, H = 256, 3
U = rng.uniform(-0.3, 0.3, (B, H, m)); x0 = rng.uniform(-0.5, 0.5, E)
br.cost_and_grad(x0, U, -1.0, host_out=True)
t0 = time.perf_counter(); c, gr = br.cost_and_grad(x0, U, -1.0, host_out=True); t1 = time.perf_counter()
print(f"B={B} H={H}: {t1 - t0:.3f} s -> {B / (t1 - t0) * H / 20:.1f} evals/s at H=20 equivalent; cost[0]={c[0]:.6f}")
