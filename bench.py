"""Benchmark of the GP-MPC rollout hot path (BASELINE.json metric: moment-matched rollout cost+gradient
evaluations per second, n=4096 training points, H=30).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[2]): synthetic contracting dynamics, n=4096, E=4, m=1
(D=5), H=30, B=1024 multi-start control sequences from one initial state, gamma=-1, lambda=2, sigma_f=1,
sigma_n=0.1, Q=2I, R=0.01I.  One "step" = one cost+gradient evaluation of the whole batch (B rollouts x H
horizon steps x E outputs).  With N GPUs the B rollouts are split into contiguous shards (GP replicated) and
one NCCL all-gather returns cost and gradient: total work is fixed, so scaling is "strong".

`--impl reference` times the reference's CPU algorithm (oracle/ref_port.py: the reference's own torch
operation sequence incl. the n^3 mm+trace and autograd; the Python reference itself cannot travel to the GPU
box) on the host cores, on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "gp_mpc_rollout_cost_grad_evals_per_sec"
UNIT = "evals/s"


def pair_ops_per_eval(D, E, m, H, first_step_const=True):
    """FP64-pipe instructions per (rollout, pair) of one H-step evaluation for outputs sharing one exp:
    chain (D adds, D squares, D-1 adds) + 11 (table exp) per pair; per output w, a row-sum add, half a column-sum add,
    N2_k for the state dimensions, and -- N1 from row / column sums on the 2 x 2 micro-tile -- (D - k1) / 2 column FMAs plus
    (1 + D - k1) / 32 row FMAs, k1 = first dimension whose N1 is kept.  The first step keeps only N1 of the action
    dimensions and no N2 (x0, Sigma_0 constant).  The per-column z_j (D / 2 per pair) is not counted."""
    chain = 3 * D - 1 + 11

    def per_output(k1, k2):
        return 1 + 1 + 0.5 + (D - k1) / 2.0 + k2 + (1 + D - k1) / 32.0
    full = chain + E * per_output(0, D - m)
    first = chain + E * per_output(D - m, 0) if first_step_const else full
    return first + (H - 1) * full if H >= 1 else 0


def ncu_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel`, taken from the committed ncu capture that
    profiles/ncu_traffic.json points at (never a constant typed into this file)."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))[kernel]
        return t["dram_bytes_per_launch"], t["source"]
    except Exception:
        return None, None


def synth(n, E, m, seed=0):
    rng = np.random.default_rng(seed)
    D = E + m
    S = rng.uniform(-1, 1, (n, E)); A = rng.uniform(-1, 1, (n, m))
    W = rng.normal(0, 0.3, (D, E))
    nxt = 0.9 * S + 0.2 * np.tanh(np.concatenate([S, A], 1) @ W)
    return S, A, nxt, rng


class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.rows = []
        self.stop_flag = threading.Event()

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [p.strip() for p in out.strip().split(",")]
                if len(parts) >= 7:
                    self.rows.append(parts)
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows)
        reasons = []
        for i, name in enumerate(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")):
            if any(r[3 + i].lower().startswith("active") for r in self.rows):
                reasons.append(name)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.rows[0][1]),
                "power_w_max": max(float(r[2]) for r in self.rows), "reasons": reasons, "samples": len(self.rows)}


def workload_config(n, H, B, world, E=4, ard=False):
    """`config` of the JSON line -- identical for the `ours` and the `reference` arm."""
    return {"workload": f"config3: n={n} E=4 m=1 H={H} gamma=-1, B={B} multi-start control sequences"
                        + (" (distinct lambdas per output)" if ard else ""),
            "n": n, "H": H, "B": B, "sharding": f"rollouts/{world}",
            "l2": f"per-step Wt working set {E * n * n * 8 / 2 / 1e6:.0f} MB > 126 MB L2 (inputs larger than L2)"}


def reference_available():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    try:
        import make_ref
        return make_ref.available()
    finally:
        sys.path.pop(0)


def run_ref_runner(device, n, Hs, steps, warmup, timeout=1500):
    """The unmodified reference (oracle/_ref) in a subprocess: objective+gradient of one control sequence at the
    full n for each horizon in Hs.  Returns the runner's dict."""
    cmd = [sys.executable, os.path.join(ROOT, "oracle", "ref_runner.py"), "--device", device, "--n", str(n),
           "--H", ",".join(str(h) for h in Hs), "--steps", str(steps), "--warmup", str(warmup)]
    env = dict(os.environ)
    if int(env.get("WORLD_SIZE", "1")) > 1:
        env.pop("OMP_NUM_THREADS", None)          # torchrun pins it to 1 per rank; the reference arm gets all host cores
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT"):
        env.pop(k, None)
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=env)
    if r.returncode != 0:
        raise RuntimeError("ref_runner failed: " + r.stderr[-2000:])
    return json.loads(r.stdout.strip().splitlines()[-1])


def extrapolate_reference(res, H_full):
    """t(H) measured at H=1 and H=2 -> t(H_full) by the measured per-step increment (the cost is linear in H:
    every horizon step runs the same E x (mean_prop + variance_prop) and keeps its own autograd state)."""
    t1 = float(np.mean(res["runs"]["1"]["times_s"]))
    out = {"t_H1_s": t1}
    if "2" in res["runs"]:
        t2 = float(np.mean(res["runs"]["2"]["times_s"]))
        out.update({"t_H2_s": t2, "t_H2_over_t_H1": t2 / t1, "t_full_s": t1 + (H_full - 1) * (t2 - t1)})
    else:
        out["t_full_s"] = t1 * H_full
    return out


def cpu_reference_sample(n, E, m, H_full, H_sample, seed, threads, device="cpu"):
    """Fallback when oracle/_ref is absent: the torch port of the reference's op sequence (oracle/ref_port.py) on a
    bounded sample: full n, H_sample horizon steps, one control sequence."""
    import torch
    from oracle.ref_port import RefPortProblem
    torch.set_num_threads(threads)
    S, A, nxt, rng = synth(n, E, m, seed)
    X = np.concatenate([S, A], 1)
    t0 = time.perf_counter()
    prob = RefPortProblem(X, nxt, np.full((E, E + m), 2.0), np.ones(E), np.full(E, 0.1), -1.0, 2 * np.eye(E),
                          0.01 * np.eye(m), device=device)
    t_fit = time.perf_counter() - t0
    x0 = rng.uniform(-0.5, 0.5, E); U = rng.uniform(-0.3, 0.3, (H_sample, m))

    def one():
        t = time.perf_counter()
        prob.cost_and_grad(x0, U)
        return time.perf_counter() - t
    return prob, one, t_fit


def run_reference(args):
    """The reference arm: the reference's OWN modules (oracle/_ref, staged from /root/reference by
    oracle/make_ref.py) on the box's host cores.  One step = objective + gradient of one control sequence at the
    full n for H=1; the same is timed for H=2 and the H=30 evaluation is extrapolated with the measured per-step
    increment.  Falls back to the torch port (kind "port") only if oracle/_ref was not staged."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n, E, m, H = args.n, 4, 1, args.H
    threads = os.cpu_count() or 1
    if reference_available():
        res = run_ref_runner("cpu", n, [1, 2], args.steps, args.warmup)
        # the H=2 leg repeats the step count of the H=1 leg: both are bounded samples of the same workload
        ex = extrapolate_reference(res, H)
        kind, t, t_full = "reference", ex["t_H1_s"], ex["t_full_s"]
        sample = (f"unmodified reference (oracle/_ref: src/mpc.py objective+gradient, torch CPU fp64, {res['cores']} threads): "
                  f"ONE control sequence at full n={n}; measured H=1 {ex['t_H1_s']:.2f} s and H=2 {ex.get('t_H2_s', float('nan')):.2f} s "
                  f"(ratio {ex.get('t_H2_over_t_H1', float('nan')):.2f}), H={H} extrapolated with the measured per-step increment "
                  f"to {t_full:.1f} s; fit ({res['fit_s']:.1f} s) excluded")
        extra = {"linearity": ex, "cores": res["cores"]}
        threads = res["cores"]
    else:
        _, one, t_fit = cpu_reference_sample(n, E, m, H, 1, 0, threads)
        for _ in range(args.warmup):
            one()
        t = float(np.mean([one() for _ in range(args.steps)]))
        t_full = t * H
        kind = "port"
        sample = (f"oracle/_ref not staged: torch-CPU port of the reference op sequence (oracle/ref_port.py), ONE control "
                  f"sequence at full n={n}, H=1, extrapolated linearly to H={H}; fit ({t_fit:.1f} s) excluded")
        extra = {}
    value = 1.0 / t_full
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(n, H, args.B, args.gpus),
        "cpu_baseline": dict({"value": value, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample}, **extra),
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_ours(args):
    import torch
    import gpmpc_b200 as gp

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    n, E, m, H, B = args.n, 4, 1, args.H, args.B
    D = E + m
    S, A, nxt, rng = synth(n, E, m, 0)
    dyn = gp.Dynamics(E, m)
    for a in range(E):
        dyn.gpr_err[a].set_lambdas(np.full(D, 2.0 + (0.1 * a if args.ard else 0.0))); dyn.gpr_err[a].set_sigma_n(np.float64(0.1))
    t0 = time.perf_counter()
    dyn.append_train_data(S, A, nxt)
    dyn._bundle.synchronize()
    t_fit = time.perf_counter() - t0
    Q = 2 * np.eye(E); R = 0.01 * np.eye(m)
    br = gp.BatchedRollouts(dyn, Q, R)
    bundle = dyn._bundle
    dev = torch.device("cuda", local)

    x0 = rng.uniform(-0.5, 0.5, E)
    U_all = rng.uniform(-0.3, 0.3, (B, H, m))
    lo, hi = gp.shard_range(B, world, rank)
    Bl = hi - lo
    per = (B + world - 1) // world
    # host (pinned) and device copies of this rank's shard
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()   # noqa: E731
    x0_h = pin(np.broadcast_to(x0, (Bl, E)).copy()); U_h = pin(U_all[lo:hi]); g_h = pin(np.full(Bl, -1.0))
    x0_d, U_d, g_d = x0_h.to(dev), U_h.to(dev), g_h.to(dev)
    cost_h = torch.empty(Bl, dtype=torch.float64).pin_memory()
    grad_h = torch.empty((Bl, H, m), dtype=torch.float64).pin_memory()
    packed = torch.zeros((per, 1 + H * m), dtype=torch.float64, device=dev)
    gathered = torch.empty((world * per, 1 + H * m), dtype=torch.float64, device=dev)
    full_h = torch.empty((world * per, 1 + H * m), dtype=torch.float64).pin_memory()

    def step_device():
        cost, grad, _, _ = bundle.cost_grad(x0_d, U_d, g_d, Q, R, want_grad=True, host_out=False)
        if world > 1:
            packed[:Bl, 0] = cost
            packed[:Bl, 1:] = grad.reshape(Bl, H * m)
            dist.all_gather_into_tensor(gathered, packed)
        return cost, grad

    def step_e2e():
        # public API with HOST buffers: H2D of the inputs and D2H of cost/grad happen inside the C-ABI call
        cost, grad, _, _ = bundle.cost_grad(x0_h.numpy(), U_h.numpy(), g_h.numpy(), Q, R, want_grad=True, host_out=True)
        if world > 1:
            packed[:Bl, 0] = torch.from_numpy(cost).to(dev, non_blocking=True)
            packed[:Bl, 1:] = torch.from_numpy(grad.reshape(Bl, H * m)).to(dev, non_blocking=True)
            dist.all_gather_into_tensor(gathered, packed)
            full_h.copy_(gathered, non_blocking=True)
            torch.cuda.synchronize()
        return cost, grad

    def timed(fn, K, sample_clocks=False):
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        sampler = None
        if sample_clocks and rank == 0:
            sampler = ClockSampler(local); sampler.start()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        l0 = bundle.launch_count()
        e0.record()
        for _ in range(K):
            fn()
        e1.record()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        ms = e0.elapsed_time(e1)
        launches = bundle.launch_count() - l0
        if sampler is not None:
            sampler.stop_flag.set(); sampler.join(2)
        if dist is not None:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, launches, (sampler.summary() if sampler is not None else None)

    for _ in range(args.warmup):
        step_device()
    ms, launches, clocks = timed(step_device, args.steps, sample_clocks=True)
    value = B * args.steps / (ms * 1e-3)

    for _ in range(max(1, args.warmup // 2)):
        step_e2e()
    ms_e2e, _, _ = timed(step_e2e, args.steps)
    e2e_value = B * args.steps / (ms_e2e * 1e-3)

    # dominant kernel: the pair-sum kernel, timed with CUDA events on its own stream inside the library
    bundle.pair_kernel_timing()              # arms the timers
    step_device(); torch.cuda.synchronize()
    pair_ms, pair_evals = bundle.pair_kernel_timing()     # summed over the H launches of one evaluation
    bundle.set_pair_timing(False)                          # armed timers synchronise after every horizon step
    fma_tflops, exp_gops = bundle.measure_fp64_peak()
    pairs = pair_evals / E                                 # (rollout, pair) evaluations, 4 outputs each
    pairs_per_rollout_step = n * (n + 1) / 2
    assert abs(pairs - Bl * H * pairs_per_rollout_step) < 1e-6 * pairs
    # --ard: one launch per output (no shared exp), same per-output accumulation
    ops = pair_ops_per_eval(D, E, m, H) if not args.ard else E * pair_ops_per_eval(D, 1, m, H)
    achieved_tflops = Bl * pairs_per_rollout_step * ops * 2.0 / (pair_ms * 1e-3) / 1e12
    bytes_algo = H * E * (n * (n + 1) / 2) * 8.0           # Wt upper triangle once per step and output
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
    nominal_tflops = sms * 64 * 2 * sm_mhz * 1e6 / 1e12      # 64 FP64 FMA / clk / SM at the sampled SM clock
    traffic, traffic_src = ncu_traffic("mm_pairs_batch")
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    roofline = {
        "bound": "fp64_pipe", "kernel": "mm_pairs_batch<5,4,grad>", "achieved": achieved_tflops, "peak": fma_tflops,
        "unit": "TFLOP/s", "frac": achieved_tflops / fma_tflops if fma_tflops else None,
        "peak_nominal": nominal_tflops, "frac_nominal": achieved_tflops / nominal_tflops,
        "traffic": traffic, "traffic_source": traffic_src,
        "fp64_instr_per_pair_per_eval": ops,
        "survey_convention": {"flop_per_pair": 36, "achieved": Bl * H * E * pairs_per_rollout_step * 36 / (pair_ms * 1e-3) / 1e12,
                              "unit": "TFLOP/s", "note": "SURVEY 8(d): P = H E n(n+1)/2 pairs x 36 flop (exp counted as 1)"},
        "note": "executed FP64-pipe instructions (x2 flop) per launch / CUDA-event duration of the pair kernel; `peak` = DFMA "
                "rate measured live by gpmpc_measure_fp64_peak (MEASURED_PEAKS.json has no fp64 figure), `peak_nominal` = "
                "SMs x 64 FMA x 2 x sampled SM clock; this kernel is neither HBM- nor tensor-bound (see DESIGN.md)",
        "pair_kernel_ms_per_eval": pair_ms, "pair_kernel_share_of_step": pair_ms / (ms / args.steps),
        "launches_per_eval": H, "fp64_exp_gops_measured": exp_gops,
        "hbm": {"achieved": bytes_algo / (pair_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                "frac": bytes_algo / (pair_ms * 1e-3) / 1e9 / hbm_peak,
                "peak_source": "MEASURED_PEAKS.json" if "hbm_gbs" in peaks else "fallback"},
    }

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": workload_config(n, H, B, world, E, args.ard), "fit_s_cold": t_fit,
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e / args.steps,
                "h2d_bytes_per_step": int(B * (E + H * m + 1) * 8), "d2h_bytes_per_step": int(B * (1 + H * m) * 8)},
        "gpu_launches": int(launches),
        "roofline": roofline,
    }
    def single_solve_section():
        # single-sequence latency (the cyipopt callback pattern, B = 1) and the p50 of one NLP solve on the same GP
        mpc = gp.RiskSensitiveMPC(-1.0, H, E, m, Q, R)
        mpc.dynamics = dyn
        mpc.set_lb([-1.0] * m); mpc.set_ub([1.0] * m)
        mpc.curr_state = torch.tensor(x0, device=dev)
        xs = U_all[0].reshape(-1)
        for _ in range(3):
            mpc.objective(xs); mpc.gradient(xs)
        lat = []
        for i in range(10):
            xi = xs + 1e-3 * (i + 1)
            t1 = time.perf_counter(); mpc.objective(xi); mpc.gradient(xi); lat.append(time.perf_counter() - t1)
        solves = []
        for i in range(5):
            n0 = mpc.n_evals
            t1 = time.perf_counter(); mpc.get_optimal_trajectory(rng.uniform(-0.5, 0.5, E)); dt = time.perf_counter() - t1
            solves.append((dt, mpc.n_evals - n0))
        try:
            import cyipopt  # noqa: F401
            solver = "cyipopt"
        except ImportError:
            solver = "scipy L-BFGS-B over the same callbacks (cyipopt is not installed)"
        # the B = 1 step kernel streams the Wt upper triangles once per horizon step: it is the HBM-bound kernel of the path
        bundle.pair_kernel_timing()
        x1 = torch.tensor(x0[None, :], device=dev); U1 = torch.tensor(U_all[:1], device=dev); g1 = torch.tensor([-1.0], device=dev, dtype=torch.float64)
        bundle.cost_grad(x1, U1, g1, Q, R, want_grad=True, host_out=False); torch.cuda.synchronize()
        single_ms, _ = bundle.pair_kernel_timing()
        bundle.set_pair_timing(False)
        # the same evaluation un-instrumented (consecutive steps overlap through programmatic dependent launch): whole
        # evaluation, device buffers, adjoint and prologue included -- a lower bound of the step kernels' bandwidth
        ev = []
        for _ in range(12):
            torch.cuda.synchronize(); t1 = time.perf_counter()
            bundle.cost_grad(x1, U1, g1, Q, R, want_grad=True, host_out=False); torch.cuda.synchronize()
            ev.append(time.perf_counter() - t1)
        eval_s = float(np.median(ev[2:]))
        line["roofline_single"] = {
            "bound": "hbm", "kernel": "mm_step_single<5,4,grad> (B=1: one fused launch per horizon step)",
            "achieved": bytes_algo / (single_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
            "frac": bytes_algo / (single_ms * 1e-3) / 1e9 / hbm_peak,
            "traffic": ncu_traffic("mm_step_single")[0], "traffic_source": ncu_traffic("mm_step_single")[1],
            "ms_per_launch": single_ms / H,
            "in_evaluation": {"ms_per_evaluation": 1e3 * eval_s, "achieved": bytes_algo / eval_s / 1e9,
                              "frac": bytes_algo / eval_s / 1e9 / hbm_peak,
                              "note": "algorithmic bytes of the H steps / wall time of one whole un-instrumented B=1 "
                                      "evaluation (device buffers, prologue and adjoint included)"},
            "note": "algorithmic bytes = H*E*n(n+1)/2*8 (Wt upper triangle once per step) / CUDA-event time of the H "
                    "launches of one B=1 evaluation (events on the library stream, programmatic dependent launch off "
                    "between timed launches)"}
        line["single_solve"] = {"objective_plus_gradient_ms": 1e3 * float(np.median(lat)),
                                "solve_p50_ms_lbfgsb_surrogate": 1e3 * float(np.median([t for t, _ in solves])),
                                "evals_per_solve": [k for _, k in solves], "solver": solver,
                                "note": "B=1 path (one fused launch per horizon step), wall clock incl. host<->device copies"}

    def cpu_baseline_section():
        threads = os.cpu_count() or 1
        if reference_available():
            # the unmodified reference (oracle/_ref) on the host cores, bounded sample: H=1 and H=2 at the full n
            res = run_ref_runner("cpu", n, [1, 2], 2, 1)
            ex = extrapolate_reference(res, H)
            line["cpu_baseline"] = {
                "value": 1.0 / ex["t_full_s"], "unit": UNIT, "cores": res["cores"], "kind": "reference", "linearity": ex,
                "sample": f"unmodified reference (oracle/_ref, src/mpc.py objective+gradient, torch CPU fp64): ONE control "
                          f"sequence at n={n}; H=1 {ex['t_H1_s']:.2f} s, H=2 {ex['t_H2_s']:.2f} s, H={H} extrapolated with the "
                          f"measured per-step increment to {ex['t_full_s']:.1f} s"}
            try:
                # context only: the same unmodified reference on THIS GPU (it picks cuda:0 when it sees one, src/gpr.py:22)
                resg = run_ref_runner("cuda", n, [1, 2], 3, 3)      # 3 warm-up calls: cuBLAS / cuSOLVER initialisation
                exg = extrapolate_reference(resg, H)
                line["reference_on_this_gpu"] = {
                    "value": 1.0 / exg["t_full_s"], "unit": UNIT, "kind": "reference", "linearity": exg,
                    "sample": f"unmodified reference (oracle/_ref) with its own device choice cuda:0 (eager torch fp64, explicit "
                              f"inverse, n^3 mm + trace, autograd): one control sequence at n={n}; H=1 {exg['t_H1_s']:.3f} s, "
                              f"H=2 {exg['t_H2_s']:.3f} s, extrapolated to H={H}"}
            except Exception as ex2:
                line["reference_on_this_gpu"] = {"error": repr(ex2)[-500:]}
        else:
            _, one, tf = cpu_reference_sample(n, E, m, H, 1, 0, threads)
            one()
            t = min(one(), one())
            line["cpu_baseline"] = {
                "value": 1.0 / (t * H), "unit": UNIT, "cores": threads, "kind": "port",
                "sample": f"oracle/_ref not staged; oracle/ref_port.py (reference op sequence, torch CPU fp64): objective+gradient "
                          f"of ONE control sequence at n={n}, H=1 ({t:.2f} s), extrapolated linearly to H={H}"}
        try:
            from oracle import oracle as orc
            orc.c_set_threads(threads)
            X = np.concatenate([S, A], 1)
            sub = 1024                                      # O(n^2) C restatement on an n=1024 sub-sample
            lam = np.full((E, D), 2.0)
            fits = [orc.fit(X[:sub], nxt[:sub, a], lam[a], 1.0, 0.1) for a in range(E)]
            t1 = time.perf_counter()
            orc.c_rollout_cost_grad(X[:sub], [f["Ky_inv"] for f in fits], [f["beta"] for f in fits], lam, np.ones(E),
                                    x0, U_all[0, :2], -1.0, Q, R)
            tc = (time.perf_counter() - t1) * (n / sub) ** 2 * (H / 2)
            line["cpu_baseline_c_oracle"] = {"value": 1.0 / tc, "unit": UNIT, "cores": threads, "kind": "port",
                                             "sample": f"oracle/gpmpc_oracle.c (O(n^2) pair sums, OpenMP): n={sub}, H=2 "
                                                       f"scaled by (n/{sub})^2 * H/2"}
        except Exception as ex:                             # the C oracle is optional here
            line["cpu_baseline_c_oracle"] = {"error": str(ex)}

    # secondary numbers must never cost the headline line
    if rank == 0 and world == 1 and not args.no_latency:
        try:
            single_solve_section()
        except Exception as ex:
            line["single_solve"] = {"error": repr(ex)}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            cpu_baseline_section()
        except Exception as ex:
            line["cpu_baseline"] = {"error": repr(ex)}
    if rank == 0 and world == 1 and not args.no_extras:
        # the other BASELINE configurations, in a subprocess: nothing that happens there can cost the headline line
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--section", "extras"], capture_output=True, text=True,
                               timeout=900)
            ex = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])
            line["configs"] = {k: ex[k] for k in ("1", "2", "4", "5") if k in ex}
            line["fit"] = ex.get("fit")
        except Exception as ex:                  # noqa: BLE001
            line["configs"] = {"error": repr(ex)[-400:]}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def make_dynamics(gp, n, E, m, ard=False, seed=0):
    S, A, nxt, rng = synth(n, E, m, seed)
    dyn = gp.Dynamics(E, m)
    for a in range(E):
        dyn.gpr_err[a].set_lambdas(np.full(E + m, 2.0 + (0.1 * a if ard else 0.0))); dyn.gpr_err[a].set_sigma_n(np.float64(0.1))
    t0 = time.perf_counter()
    dyn.append_train_data(S, A, nxt)
    dyn._bundle.synchronize()
    return dyn, (S, A, nxt), rng, time.perf_counter() - t0


def run_config5(args):
    """BASELINE configs[4]: a gamma sweep x initial states, every (gamma, x0) an independent MPC problem (n=4096, H=30),
    solved in lock step per GPU by BatchedSolver and partitioned over the ranks (GP replicated, one all-gather of the
    solutions).  Prints one JSON line: solves/s (whole job) and the rollout evaluations/s behind it."""
    import torch
    import gpmpc_b200 as gp
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    n, E, m, H = args.n, 4, 1, args.H
    dyn, _, rng, t_fit = make_dynamics(gp, n, E, m)
    Q = 2 * np.eye(E); R = 0.01 * np.eye(m)
    gammas = np.array([-2.0, -1.0, 0.5, 1.0])
    n_x0 = args.instances // len(gammas)
    starts = rng.uniform(-0.5, 0.5, (n_x0, E))
    G, I = np.meshgrid(gammas, np.arange(n_x0), indexing="ij")
    gam = G.reshape(-1); x0 = starts[I.reshape(-1)]
    B = gam.size
    br = gp.BatchedRollouts(dyn, Q, R)
    # warm-up: one batched evaluation of this rank's shard size (kernel selection, workspaces)
    nb = len(gp.shard_indices(B, world, rank))
    br.cost_and_grad(x0[:nb], np.zeros((nb, H, m)), gam[:nb], host_out=True)
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    solver = gp.BatchedSolver(br, H, m, lb=[-1.0], ub=[1.0], max_iter=40, gtol=1e-4)
    if world > 1:
        sol = solver.solve_sharded(x0, gam)
    else:
        r = solver.solve(x0, gam)
        sol = {"U": r["U"], "cost": r["cost"], "converged": r["converged"], "iters_per_rank": np.array([r["iters"]]),
               "rollout_evals": r["rollout_evals"]}
    e1.record()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    ms = e0.elapsed_time(e1)
    if dist is not None:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    if rank == 0:
        print(json.dumps({
            "metric": "gp_mpc_closed_form_solves_per_sec", "value": B / (ms * 1e-3), "unit": "solves/s", "n_gpus": world,
            "ms_total": ms, "higher_is_better": True, "scaling": "strong", "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"config5: gamma sweep {gammas.tolist()} x {n_x0} initial states = {B} MPC instances, n={n} "
                                   f"E=4 m=1 H={H}, projected L-BFGS in lock step per GPU (gtol 1e-4, max 40 iterations)",
                       "sharding": f"instances interleaved over {world} rank(s), GP replicated, one all-gather of U/cost"},
            "rollout_evals_per_sec": sol["rollout_evals"] / (ms * 1e-3),
            "rollout_evals_per_instance": sol["rollout_evals"] / B,
            "converged_fraction": float(np.mean(sol["converged"])), "iterations_per_rank": sol["iters_per_rank"].tolist(),
            "fit_s": t_fit}), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def run_split(args):
    """ONE control sequence (B = 1, the IPOPT callback) whose every horizon step is split over the GPUs of the node:
    each rank sweeps 1/N of the pair space in the persistent kernel and the kernels exchange their per-step sums over
    NVLink (peer-mapped mailboxes, no NCCL call per step).  Prints one JSON line: ms per objective+gradient evaluation."""
    import torch
    import gpmpc_b200 as gp
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    n, E, m, H = args.n, 4, 1, args.H
    dyn, _, rng, t_fit = make_dynamics(gp, n, E, m)
    Q = 2 * np.eye(E); R = 0.01 * np.eye(m)
    bundle = dyn._bundle
    x0 = torch.tensor(rng.uniform(-0.5, 0.5, (1, E)), device=dev); U = torch.tensor(rng.uniform(-0.3, 0.3, (1, H, m)), device=dev)
    g = torch.full((1,), -1.0, dtype=torch.float64, device=dev)
    ref_cost, ref_grad, _, _ = bundle.cost_grad(x0, U, g, Q, R, host_out=True)       # single-GPU result (every rank, identical)
    if world > 1:
        bundle.split_connect()
    def one():
        return bundle.cost_grad(x0, U, g, Q, R, host_out=False)
    for _ in range(args.warmup):
        one()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        cost, grad, _, _ = one()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    exch = (0.0, 0.0)
    if world > 1:
        bundle.set_option("split_timeline", 1)
        one(); torch.cuda.synchronize()
        exch = bundle.split_last_exchange_us()
        bundle.set_option("split_timeline", 0)
        t = torch.tensor([ms, exch[0], exch[1]], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, exch = float(t[0]), (float(t[1]), float(t[2]))
    err_c = abs(float(cost[0]) - float(ref_cost[0])) / max(1.0, abs(float(ref_cost[0])))
    err_g = float(np.max(np.abs(grad.cpu().numpy() - ref_grad))) / max(1e-300, float(np.max(np.abs(ref_grad))))
    bytes_algo = H * E * (n * (n + 1) / 2) * 8.0
    if rank == 0:
        print(json.dumps({
            "metric": "gp_mpc_single_rollout_cost_grad_ms", "value": ms, "unit": "ms", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "higher_is_better": False, "scaling": "strong", "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"one control sequence (B=1), n={n} E=4 m=1 H={H}: objective + gradient, every step's pair space "
                                   f"split over {world} GPU(s)"},
            "us_per_horizon_step": 1e3 * ms / H, "aggregate_wt_stream_GBs": bytes_algo / (ms * 1e-3) / 1e9,
            "exchange_us_per_step_mean": exch[0], "exchange_us_per_step_max": exch[1],
            "vs_single_gpu_result": {"cost_rel_err": err_c, "grad_err_over_max": err_g}, "fit_s": t_fit}), flush=True)
    if dist is not None:
        bundle.split_disconnect()
        dist.destroy_process_group()


def run_extras(args):
    """The other BASELINE configurations and the "next" rows on ONE GPU (run as a subprocess of the default bench so that
    a failure here can never cost the headline line).  Prints one JSON dict."""
    import torch
    import gpmpc_b200 as gp
    torch.cuda.set_device(0)
    out = {}

    def timed(fn, reps=2):
        fn(); torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
        return min(ts)

    def section(name, fn):
        try:
            out[name] = fn()
        except Exception as ex:                  # noqa: BLE001
            out[name] = {"error": repr(ex)[-400:]}
        torch.cuda.empty_cache()

    def config1():
        g = np.load(os.path.join(ROOT, "tests", "golden", "shipped.npz"))
        mpc = gp.RiskSensitiveMPC(-1, 6, 2, 2, 2 * np.identity(2), np.zeros((2, 2)), None)
        for i in range(2):
            mpc.dynamics.gpr_err[i].set_sigma_n(np.float64(g["ship_sn"][i]))
            mpc.dynamics.gpr_err[i].set_lambdas(np.asarray(g["ship_lam"][i], dtype=np.float64))
            mpc.dynamics.gpr_err[i].set_sigma_f(np.float64(g["ship_sf"][i]))
        mpc.dynamics.append_train_data(g["ship_S"], g["ship_A"], g["ship_next"])
        mpc.set_xref(np.array([0., 0.])); mpc.set_uref(np.array([0., 0.])); mpc.set_lb([-1.0, -1.0]); mpc.set_ub([1.0, 1.0])
        mpc.curr_state = torch.tensor(g["ship_x0"], device="cuda")
        x = g["ship_U0"].reshape(-1).copy()
        for _ in range(3):
            mpc.objective(x); mpc.gradient(x)
        ts = []
        for i in range(50):
            xi = x + 1e-4 * i
            t0 = time.perf_counter(); mpc.objective(xi); mpc.gradient(xi); ts.append(time.perf_counter() - t0)
        st = []
        for i in range(5):
            n0 = mpc.n_evals; t0 = time.perf_counter(); mpc.get_optimal_trajectory(g["ship_x0"]); st.append((time.perf_counter() - t0, mpc.n_evals - n0))
        return {"workload": "the reference's own experiment (shipped data): n=400 E=2 m=2 H=6 gamma=-1",
                "objective_plus_gradient_ms": 1e3 * float(np.median(ts)),
                "solve_p50_ms_lbfgsb_surrogate": 1e3 * float(np.median([t for t, _ in st])), "evals_per_solve": [k for _, k in st]}

    def config2():
        dyn, _, rng, tf = make_dynamics(gp, 2048, 4, 1)
        U = torch.tensor(rng.uniform(-0.5, 0.5, (8192, 5)), device="cuda"); Sd = torch.tensor(rng.uniform(1e-3, 5e-2, (8192, 5)), device="cuda")
        t = timed(lambda: dyn._bundle.moment_match(U, Sd, out_device=True), 3)
        return {"workload": "n=2048 D=5 E=4, 8192 uncertain inputs (mean + variance)", "ms_per_batch": 1e3 * t,
                "inputs_per_sec": 8192 / t, "fit_s_cold": tf}

    def config4():
        dyn, _, rng, tf = make_dynamics(gp, 16384, 4, 1)
        t0 = time.perf_counter(); dyn._fit_all(); dyn._bundle.synchronize(); t_fit = time.perf_counter() - t0
        Q = 2 * np.eye(4); R = 0.01 * np.eye(1)
        H = 20
        res = {"workload": "n=16384 E=4 m=1 H=20", "fit_s_warm": t_fit, "fit_s_cold": tf}
        for B, full in ((128, True), (256, False)):
            U = torch.tensor(rng.uniform(-0.3, 0.3, (B, H, 1)), device="cuda"); x0 = torch.tensor(rng.uniform(-0.5, 0.5, (B, 4)), device="cuda")
            gm = torch.full((B,), -1.0, dtype=torch.float64, device="cuda")
            t = timed(lambda: dyn._bundle.cost_grad(x0, U, gm, Q, R, host_out=False, full=full), 1)
            key = "full_covariance" if full else "variance_only"
            res[key] = {"B": B, "s_per_eval_batch": t, "evals_per_sec": B / t}
            if full:
                # FP64-pipe instructions per (rollout, pair) and step, D=5, 10 pair-outputs sharing one exp:
                # forward 14 + 11 + 10, backward 14 + 11 + 10 + 1 + 5 + 5 + 15 (the z_i / z_j transforms are extra)
                ops = (35 + 61) * H
                pairs = 16384 * 16385 / 2
                tf64 = B * pairs * ops * 2 / t / 1e12
                peak, _ = dyn._bundle.measure_fp64_peak()
                res[key].update({"fp64_tflops": tf64, "fp64_peak_measured": peak, "fp64_frac": tf64 / peak,
                                 "fp64_frac_nominal": tf64 / 37.2, "fp64_instr_per_pair_step": 96})
        return res

    def config5():
        dyn, _, rng, tf = make_dynamics(gp, 4096, 4, 1)
        br = gp.BatchedRollouts(dyn, 2 * np.eye(4), 0.01 * np.eye(1))
        gammas = np.array([-2.0, -1.0, 0.5, 1.0]); starts = rng.uniform(-0.5, 0.5, (128, 4))
        G, I = np.meshgrid(gammas, np.arange(128), indexing="ij")
        solver = gp.BatchedSolver(br, 30, 1, lb=[-1.0], ub=[1.0], max_iter=40, gtol=1e-4)
        t0 = time.perf_counter(); sol = solver.solve(starts[I.reshape(-1)], G.reshape(-1)); t = time.perf_counter() - t0
        return {"workload": "512 MPC instances (4 gammas x 128 x0), n=4096 H=30, lock-step projected L-BFGS, one GPU",
                "solves_per_sec": G.size / t, "s_total": t, "rollout_evals_per_instance": sol["rollout_evals"] / G.size,
                "converged_fraction": float(sol["converged"].mean())}

    def fit_n2_n3():
        res = {}
        for ard in (False, True):
            dyn, _, rng, tf = make_dynamics(gp, 4096, 4, 1, ard=ard)
            ts = []
            for _ in range(3):
                t0 = time.perf_counter(); dyn._fit_all(); dyn._bundle.synchronize(); ts.append(time.perf_counter() - t0)
            key = "distinct_hypers" if ard else "shared_hypers"
            # DMMA work counted as n^3 per factorisation (Cholesky n^3/3 + triangular inverse n^3/3 + Z^T Z n^3/3)
            nfac = 4 if ard else 1
            res[key] = {"n": 4096, "outputs": 4, "warm_ms": 1e3 * min(ts), "factorisations": nfac,
                        "dmma_frac_of_nominal": nfac * 4096.0 ** 3 / min(ts) / 37.2e12}
        ts = []
        for _ in range(5):
            s = rng.uniform(-1, 1, 4); a = rng.uniform(-1, 1, 1)
            t0 = time.perf_counter(); dyn.append_train_data(s, a, 0.9 * s); dyn._bundle.synchronize(); ts.append(time.perf_counter() - t0)
        res["n2_append_ms"] = 1e3 * float(np.median(ts))
        g = dyn.gpr_err[0]
        t0 = time.perf_counter(); g.update_hyperparams(num_iters=5, verbose=False); dyn._bundle.synchronize()
        res["n3_adam_step_ms"] = 1e3 * (time.perf_counter() - t0) / 5
        return res

    section("1", config1); section("2", config2); section("4", config4); section("5", config5); section("fit", fit_n2_n3)
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", "--ntrain", dest="n", type=int, default=4096,
                    help="training points (use --ntrain under torchrun: its own parser claims the abbreviation --n)")
    ap.add_argument("--H", type=int, default=30)
    ap.add_argument("--B", type=int, default=1024)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-latency", action="store_true")
    ap.add_argument("--ard", action="store_true", help="distinct length-scales per output (not the headline workload)")
    ap.add_argument("--no-extras", action="store_true", help="skip the other BASELINE configurations (configs 1, 2, 4, 5, fit)")
    ap.add_argument("--section", default="", choices=["", "extras"], help="internal: run one secondary section and print its JSON")
    ap.add_argument("--config", type=int, default=3, choices=[3, 5, 7],
                    help="3 = headline rollout benchmark, 5 = sharded MPC solves, 7 = one rollout split over the GPUs (latency)")
    ap.add_argument("--instances", type=int, default=2048, help="--config 5: number of MPC instances (4 gammas x instances/4 x0)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.section == "extras":
        run_extras(args)
    elif args.config == 5:
        run_config5(args)
    elif args.config == 7:
        run_split(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
