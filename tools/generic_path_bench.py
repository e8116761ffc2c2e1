import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import time, torch, numpy as np
import optimistix_b200 as optx
from optimistix_b200 import synthetic
DEV="cuda:0"
B=65536
w = synthetic.exp_decay_batch(B)
t, d, y0 = (torch.as_tensor(w[k], device=DEV) for k in ("t", "d", "y0"))
def fn(y, a):
    tt, dd = a
    return y[0] * torch.exp(-y[1] * tt) + y[2] - dd
def run_gen():
    return optx.least_squares(fn, optx.LevenbergMarquardt(1e-8, 1e-8), y0, (t, d), args_in_dims=(None, 0), throw=False)
def run_per():
    return optx.least_squares(optx.models.exp_decay, optx.LevenbergMarquardt(1e-8, 1e-8), y0, (t, d), throw=False)
import os
for name, f in (("generic", run_gen), ("persistent", run_per)):
    f(); torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        t0 = time.perf_counter(); s = f(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    dt = min(ts)
    print(name, "graph=" + os.environ.get("B200OPT_GENERIC_GRAPH", "1"), " ".join(f"{x*1e3:.1f}" for x in ts), "ms ->",
          f"{B/dt/1e6:.2f} M problems/s", "launches", s.stats.get("step_kernel_launches"))

# long iteration counts: BFGS on the chained Rosenbrock function (n = 64, ~760 iterations per problem)
yr = torch.as_tensor(synthetic.rosenbrock_batch(256)["y0"], device=DEV)
def rosen(y, _):
    return (100.0 * (y[1:] - y[:-1] ** 2) ** 2 + (1.0 - y[:-1]) ** 2).sum()
def run_bfgs():
    return optx.minimise(rosen, optx.BFGS(1e-8, 1e-8), yr, None, max_steps=4096, throw=False)
run_bfgs(); torch.cuda.synchronize()
ts = []
for _ in range(3):
    t0 = time.perf_counter(); s = run_bfgs(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
print("bfgs generic graph=" + os.environ.get("B200OPT_GENERIC_GRAPH", "1"), " ".join(f"{x*1e3:.0f}" for x in ts), "ms; launches",
      s.stats["step_kernel_launches"], "ok", int((s.result == 0).sum()))
