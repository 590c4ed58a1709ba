"""Where the fixed cost of a single-control-sequence evaluation goes (development tool):
python tools/b1_breakdown.py [n] [H] [reps]   -- host vs device buffers, with / without gradient."""
import sys, os, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpmpc_b200 as gp
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
H = int(sys.argv[2]) if len(sys.argv) > 2 else 30
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 50
E, m = 4, 1
rng = np.random.default_rng(0)
S = rng.uniform(-1, 1, (n, E)); A = rng.uniform(-1, 1, (n, m))
nxt = 0.9 * S + 0.2 * np.tanh(np.concatenate([S, A], 1) @ rng.normal(0, 0.3, (E + m, E)))
dyn = gp.Dynamics(E, m)
for a in range(E):
    dyn.gpr_err[a].set_lambdas(np.full(E + m, 2.0)); dyn.gpr_err[a].set_sigma_n(np.float64(0.1))
dyn.append_train_data(S, A, nxt)
dyn._sync_propagation_hypers()
bundle = dyn._bundle
if os.environ.get('BIG'):
    bundle.set_option('single_big_share', int(os.environ['BIG'])); print('single_big_share', os.environ['BIG'])
if os.environ.get('PERSIST'):
    bundle.set_option('persistent_single', 1); print('persistent whole-horizon kernel ON')
Q = 2 * np.eye(E); R = 0.01 * np.eye(m)
x0 = np.zeros((1, E)); U = rng.uniform(-0.3, 0.3, (1, H, m)); g = np.full(1, -1.0)
dev = bundle.device
x0d, Ud, gd = (torch.from_numpy(v).to(dev) for v in (x0, U, g))

def med(f):
    for _ in range(5): f()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); f(); ts.append(time.perf_counter() - t0)
    return 1e6 * float(np.median(ts))

def dev_call(want_grad=True):
    bundle.cost_grad(x0d, Ud, gd, Q, R, want_grad=want_grad, host_out=False); torch.cuda.synchronize()

c0, g0, _, _ = bundle.cost_grad(x0, U, g, Q, R)
print(f"n={n} H={H}  cost {c0[0]:.15g} |grad| {float(np.linalg.norm(g0)):.15g}  us per evaluation (median of {reps})")
print("  host in/out, cost+grad      ", round(med(lambda: bundle.cost_grad(x0, U, g, Q, R)), 1))
print("  host in/out, cost only      ", round(med(lambda: bundle.cost_grad(x0, U, g, Q, R, want_grad=False)), 1))
print("  device in/out + sync, +grad ", round(med(dev_call), 1))
print("  device in/out + sync, cost  ", round(med(lambda: dev_call(False)), 1))
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
st = torch.cuda.current_stream()
def ev_time(want_grad):
    ts = []
    for _ in range(reps):
        ev0.record(st); bundle.cost_grad(x0d, Ud, gd, Q, R, want_grad=want_grad, host_out=False); ev1.record(st)
        torch.cuda.synchronize(); ts.append(ev0.elapsed_time(ev1) * 1e3)
    return round(float(np.median(ts)), 1)
print("  device time by events, +grad", ev_time(True), " cost only", ev_time(False))
