"""Times the UNMODIFIED reference (`oracle/_ref`, staged by oracle/make_ref.py) on the benchmark workload --
BASELINE INFRASTRUCTURE ONLY; run as a subprocess by `bench.py` (the `--impl reference` arm and the `cpu_baseline`
leg), never imported by the product package.

    python oracle/ref_runner.py --device cpu --n 4096 --H 1,2 --steps 3 --warmup 1

One step = `RiskSensitiveMPC.objective(x)` followed by `RiskSensitiveMPC.gradient(x)` (`src/mpc.py:202-255`: rollout
through `Dynamics.forward_propagate_torch`, `cost_torch`, autograd backward) for ONE control sequence at the full n.
The reference needs ~1.75 GB of autograd state per horizon step at n=4096 and ~5 s per step on 8 cores (SURVEY.md
section 0), so the full H=30 evaluation is out of reach; the runner times short horizons (default H=1 and H=2) so that
the caller can show the cost is linear in H and extrapolate.  `--device cpu` hides the GPUs before torch is imported
(the reference picks cuda:0 whenever it sees one, `src/gpr.py:22`, `src/mpc.py:41`); `--device cuda` lets it.

Prints one JSON line: {"device", "cores", "n", "fit_s", "runs": {"<H>": {"times_s": [...], "cost": c}}}.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--device", default="cpu", choices=["cpu", "cuda"])
    ap.add_argument("--n", type=int, default=4096)
    ap.add_argument("--E", type=int, default=4)
    ap.add_argument("--m", type=int, default=1)
    ap.add_argument("--H", default="1,2")
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--seed", type=int, default=0)
    a = ap.parse_args()
    if a.device == "cpu":
        os.environ["CUDA_VISIBLE_DEVICES"] = ""
    import numpy as np
    import torch
    threads = a.threads or (os.cpu_count() or 1)
    torch.set_num_threads(threads)
    for p in (ROOT, HERE):
        if p not in sys.path:
            sys.path.insert(0, p)
    from bench import synth                       # the benchmark's own generator (numpy only)
    import make_ref
    _, _, _, ref_mpc = make_ref.import_reference()

    n, E, m = a.n, a.E, a.m
    D = E + m
    S, A, nxt, rng = synth(n, E, m, a.seed)
    x0 = rng.uniform(-0.5, 0.5, E)
    Hs = [int(h) for h in a.H.split(",") if h]
    out = {"device": a.device, "cores": threads, "n": n, "runs": {}}
    for H in Hs:
        mpc = ref_mpc.RiskSensitiveMPC(-1.0, H, E, m, 2 * np.eye(E), 0.01 * np.eye(m))
        assert (mpc.device.type == "cuda") == (a.device == "cuda"), mpc.device
        for g in mpc.dynamics.gpr_err:
            g.set_lambdas(np.full(D, 2.0)); g.set_sigma_n(np.float64(0.1))
        t0 = time.perf_counter()
        mpc.dynamics.append_train_data(S, A, nxt)
        if a.device == "cuda":
            torch.cuda.synchronize()
        out.setdefault("fit_s", time.perf_counter() - t0)
        mpc.curr_state = torch.tensor(x0, device=mpc.device).type(torch.float64)
        U = np.random.default_rng(a.seed + 1).uniform(-0.3, 0.3, (a.warmup + a.steps, H * m))
        times, cost = [], None
        for i in range(a.warmup + a.steps):
            t0 = time.perf_counter()
            cost = mpc.objective(U[i])
            grad = mpc.gradient(U[i])
            if a.device == "cuda":
                torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            assert grad.shape == (H, m)
            if i >= a.warmup:
                times.append(dt)
        out["runs"][str(H)] = {"times_s": times, "cost": float(cost)}
        del mpc
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
