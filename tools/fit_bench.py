"""Warm fit time probe (development tool): python tools/fit_bench.py [n] [reps]   (GPMPC_NO_LOOKAHEAD=1 for the A/B)
Times Dynamics._fit_all (Gram, blocked Cholesky, triangular inverse, Ky^-1, beta, Wt) for 4 outputs with shared and with
distinct hyper-parameters, and checks the result against the residual |Ky^-1 Ky - I| on probe columns."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gpmpc_b200 as gp
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
E, m = 4, 1
rng = np.random.default_rng(0)
S = rng.uniform(-1, 1, (n, E)); A = rng.uniform(-1, 1, (n, m))
nxt = 0.9 * S + 0.2 * np.tanh(np.concatenate([S, A], 1) @ rng.normal(0, 0.3, (E + m, E)))
for ard in (False, True):
    dyn = gp.Dynamics(E, m)
    for a in range(E):
        dyn.gpr_err[a].set_lambdas(np.full(E + m, 2.0 + (0.1 * a if ard else 0.0))); dyn.gpr_err[a].set_sigma_n(np.float64(0.1))
    dyn.append_train_data(S, A, nxt)
    ts = []
    for _ in range(reps):
        dyn._bundle.synchronize(); t0 = time.perf_counter(); dyn._fit_all(); dyn._bundle.synchronize(); ts.append(time.perf_counter() - t0)
    g = dyn.gpr_err[E - 1]
    cols = rng.choice(n, 32, replace=False)
    Ky = g.Ky
    R = g.Ky_inv @ Ky[:, cols]
    R[cols, torch.arange(32)] -= 1.0
    print(f"n={n} {'distinct' if ard else 'shared'} hyper-parameters: warm fit {1e3 * min(ts):.2f} ms (median {1e3 * float(np.median(ts)):.2f}); "
          f"max|Ky^-1 Ky - I| on 32 columns = {float(R.abs().max()):.2e}; lookahead={'off' if os.environ.get('GPMPC_NO_LOOKAHEAD') else 'on'}")
