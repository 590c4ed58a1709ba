"""Race / determinism stress: repeated evaluations must be bit-identical (ticket counters, last-CTA reductions,
programmatic dependent launch, dynamic work items).   python tools/stress.py [repeats]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpmpc_b200 as gp

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 40
bad = 0
for n in (90, 1000, 3000):
    E, m, H = 4, 1, 8
    rng = np.random.default_rng(n)
    S = rng.uniform(-1, 1, (n, E)); A = rng.uniform(-1, 1, (n, m))
    nxt = 0.9 * S + 0.2 * np.tanh(np.concatenate([S, A], 1) @ rng.normal(0, 0.3, (E + m, E)))
    for ard in (False, True):
        dyn = gp.Dynamics(E, m)
        for a in range(E):
            dyn.gpr_err[a].set_lambdas(np.full(E + m, 2.0 + (0.2 * a if ard else 0.0))); dyn.gpr_err[a].set_sigma_n(np.float64(0.1))
        dyn.append_train_data(S, A, nxt)
        br = gp.BatchedRollouts(dyn, 2 * np.eye(E), 0.01 * np.eye(m))
        for B in (1, 3, 40, 111, 130, 300):
            x0 = rng.uniform(-0.5, 0.5, (B, E)); U = rng.uniform(-0.3, 0.3, (B, H, m))
            c0, g0 = br.cost_and_grad(x0, U, -1.0, host_out=True)
            for r in range(reps):
                # interleave other shapes so that buffers, counters and launch patterns change between repeats
                if r % 5 == 0:
                    br.cost_and_grad(x0[:1], U[:1], -1.0, host_out=True)
                c, g = br.cost_and_grad(x0, U, -1.0, host_out=True)
                if not (np.array_equal(c, c0) and np.array_equal(g, g0)):
                    bad += 1
                    print(f"MISMATCH n={n} ard={ard} B={B} repeat {r}: max |dc| {np.max(np.abs(c - c0)):.3e}")
        print(f"n={n} ard={ard}: ok so far, mismatches {bad}", flush=True)
print("stress done, mismatches:", bad)
sys.exit(1 if bad else 0)
