"""Row-sharded variant of gram_api_probe.py (torchrun, one rank per GPU): per-iteration CUDA-event times of
  (a) the step-wise kernels with the peer exchange (Gram + reduce/publish + fp64 solve waiting on the flags),
  (b) b200opt_gn_iterate called directly on a persistent state (no Python wrapper per solve),
  (c) the public API's device loop (gauss_newton_large), fp32+refinement and fp64 solver,
each `reps` times, interleaved, max over ranks.  Total rows = --rows, split evenly."""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import optimistix_b200 as optx  # noqa: E402
from optimistix_b200._gn_large import N, OUT_FLOATS, CudaOps, PeerExchange, gauss_newton_large  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=64 << 20)
ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--reps", type=int, default=4)
a = ap.parse_args()
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
m = a.rows // world // 64 * 64
gen = torch.Generator(device=dev).manual_seed(4 + rank)
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
xk = PeerExchange(dev)          # private exchange of the step-wise loop (host-numbered)
xs = PeerExchange.shared(dev)   # the exchange the device loop uses (device-numbered)
r0, r1 = torch.empty(m, device=dev), torch.empty(m, device=dev)
state = ops.state_new()


def kernels():
    for _ in range(a.iters):
        parity, seq = xk.next()
        ops.residual_gram_xchg(J, y, d, r, r_old, 1e-4, 1e-4, fmax, out, xk, parity, seq)
        ops.solve_xchg(xk, parity, seq, 0.0, p, status, summary)


def iterate_direct(solver):
    def f():
        ops.state_init(state, y0, 0.0, 0.0, a.iters, solver, 0, r1, out)
        for _ in range(a.iters):
            ops.iterate(J, d, r0, r1, state, out, xs, 0.0, 0.0)
    return f


def api(solver):
    return lambda: gauss_newton_large(J, d, y0, never, max_steps=a.iters, throw=False, linear_solver=solver, check_every=a.iters)


def timed(fn):
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / a.iters], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


cfgs = {"step-wise kernels, peer exchange, fp64 solve": kernels, "gn_iterate direct, fp64 solver": iterate_direct(1),
        "gn_iterate direct, fp32+refinement": iterate_direct(0), "public API device loop, fp64 solver": api(1),
        "public API device loop, fp32+refinement": api(0)}
for f in cfgs.values():
    f()
res = {k: [] for k in cfgs}
for _ in range(a.reps):
    for k, f in cfgs.items():
        res[k].append(timed(f))
if rank == 0:
    print(json.dumps({"world": world, "rows_per_gpu": m, "iters": a.iters,
                      "ms_per_iteration": {k: {"min": min(v), "all": [round(x, 3) for x in v]} for k, v in res.items()}}, indent=1))
dist.barrier()
xk.close()
PeerExchange.close_all()
dist.destroy_process_group()
