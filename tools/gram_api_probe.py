"""Where does an iteration of the large Gauss-Newton problem spend its time?  (run on the GPU box)
    python tools/gram_api_probe.py [--rows 16777216] [--iters 20] [--reps 5]
Interleaves, `reps` times: the step-wise kernels enqueued back to back (Gram + reduce kernel + fp64 solve: the round-1
"kernel only" loop), the device loop through gauss_newton_large with the fp32+refinement solver, and with the fp64 solver.
Prints min / median ms per iteration of each.  Under `ncu --metrics gpu__time_duration.sum` the launch list gives the
per-kernel durations of the same loops."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import optimistix_b200 as optx  # noqa: E402
from optimistix_b200._gn_large import N, OUT_FLOATS, CudaOps, gauss_newton_large  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=16 << 20)
ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--out", default="")
a = ap.parse_args()
dev = torch.device("cuda", 0)
m = a.rows
gen = torch.Generator(device=dev).manual_seed(4)
scale = torch.logspace(0, -2, N, device=dev)
J = torch.empty((m, N), device=dev)
for s0 in range(0, m, 1 << 20):
    J[s0:s0 + (1 << 20)] = torch.randn((min(1 << 20, m - s0), N), generator=gen, device=dev) * scale
d = torch.randn(m, generator=gen, device=dev)
y0 = torch.zeros(N, device=dev)
never = optx.GaussNewton(0.0, 0.0)
ops = CudaOps(dev)
y = torch.zeros(N, device=dev)
r, r_old = torch.empty(m, device=dev), torch.zeros(m, device=dev)
out = torch.empty(OUT_FLOATS, device=dev)
fmax = out[N * N + N + 1:].view(torch.int32)
summary, p, status = torch.zeros(2, device=dev), torch.empty(N, device=dev), torch.zeros(1, dtype=torch.int32, device=dev)


def timed(fn):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / a.iters


def kernels():
    for _ in range(a.iters):
        ops.residual_gram(J, y, d, r, r_old, 1e-4, 1e-4, fmax, out)
        ops.solve_parts(out, 1, OUT_FLOATS, 0.0, p, status, summary)


def gram_only():
    for _ in range(a.iters):
        ops.residual_gram(J, y, d, r, r_old, 1e-4, 1e-4, fmax, out)


def api(solver, fold=False):
    return lambda: gauss_newton_large(J, d, y0, never, max_steps=a.iters, throw=False, linear_solver=solver, check_every=a.iters, fold_reduce=fold)


def kernels_split():
    """The kernel-only loop with an event after every kernel: where the time between the big kernels goes."""
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(a.iters)]
    torch.cuda.synchronize()
    for ev in evs:
        ev[0].record()
        ops.residual_gram(J, y, d, r, r_old, 1e-4, 1e-4, fmax, out)
        ev[1].record()
        ops.solve_parts(out, 1, OUT_FLOATS, 0.0, p, status, summary)
        ev[2].record()
    torch.cuda.synchronize()
    g = sum(ev[0].elapsed_time(ev[1]) for ev in evs) / a.iters
    s = sum(ev[1].elapsed_time(ev[2]) for ev in evs) / a.iters
    gap = sum(evs[i][2].elapsed_time(evs[i + 1][0]) for i in range(a.iters - 1)) / (a.iters - 1)
    return {"gram+reduce": g, "solve": s, "gap_between_iterations": gap, "total": evs[0][0].elapsed_time(evs[-1][2]) / a.iters}


cfgs = {"gram+reduce kernels only": gram_only, "kernel-only loop (gram, reduce, fp64 solve)": kernels,
        "device loop, fp32+refinement solver": api(0), "device loop, fp64 solver": api(1),
        "device loop, fp32+refinement solver, reduction folded into the Gram kernel": api(0, True)}
for f in cfgs.values():
    f()
res = {k: [] for k in cfgs}
for _ in range(a.reps):
    for k, f in cfgs.items():
        res[k].append(timed(f))
summary_ = {k: {"min_ms": min(v), "median_ms": sorted(v)[len(v) // 2], "all": [round(x, 4) for x in v]} for k, v in res.items()}
split = [kernels_split() for _ in range(3)]
print(json.dumps({"rows": m, "iters": a.iters, "per_iteration": summary_, "kernel_only_loop_split_ms": split}, indent=1))
if a.out:
    with open(a.out, "w") as fh:
        json.dump({"rows": m, "iters": a.iters, "per_iteration": summary_}, fh, indent=1)
