"""Multi-GPU check of the public API on the row-sharded problem (run with torchrun, one rank per GPU):
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/check_gn_large_multi.py
Every rank solves its row shard through optx.least_squares(models.linear_rows, GaussNewton, ...) with the peer-memory
exchange and with the NCCL all-gather; both must give the bit-identical iterate on every rank and match fp64 lstsq."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import optimistix_b200 as optx
from optimistix_b200 import synthetic
from optimistix_b200._gn_large import N, gauss_newton_large

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
m = 262144
w = synthetic.linear_lsq_rows(m, seed=4)
lo, hi = rank * m // world, (rank + 1) * m // world
J, d = torch.as_tensor(w["J"][lo:hi], device=dev), torch.as_tensor(w["d"][lo:hi], device=dev)
y0 = torch.zeros(N, dtype=torch.float32, device=dev)
sv = optx.GaussNewton(1e-4, 1e-4)
a = optx.least_squares(optx.models.linear_rows, sv, y0, (J, d), throw=False)  # default: device loop, peer exchange, fp32 + refinement solver
assert "peer memory (device loop)" in a.stats["exchange"], a.stats
a64 = gauss_newton_large(J, d, y0, sv, throw=False, linear_solver=1)           # device loop, fp64 solver
af = gauss_newton_large(J, d, y0, sv, throw=False, linear_solver=1, fold_reduce=True)  # reduction + publication folded into the Gram kernel
assert torch.equal(af.value, a64.value) and af.stats["num_steps"] == a64.stats["num_steps"]
b = gauss_newton_large(J, d, y0, sv, throw=False, exchange="nccl")             # host loop over the NCCL all-gather (fp64 solver)
c = gauss_newton_large(J, d, y0, sv, throw=False, device_loop=False)           # host loop over the peer exchange
assert a.result == a64.result == b.result == c.result == optx.RESULTS.successful, (a.result, a64.result, b.result, c.result)
assert a64.stats["num_steps"] == b.stats["num_steps"] == c.stats["num_steps"]
assert torch.equal(a64.value, b.value) and torch.equal(b.value, c.value), "device loop, NCCL all-gather and host-loop peer exchange must give the same bits"
assert abs(a.stats["num_steps"] - b.stats["num_steps"]) <= 1 and a.stats["fp64_fallbacks"] == 0
for _ in range(3):  # repeated solves reuse the cached exchange (sequence numbers keep counting)
    a2 = optx.least_squares(optx.models.linear_rows, sv, y0, (J, d), throw=False)
    assert torch.equal(a2.value, a.value)
allv = [torch.empty_like(a.value) for _ in range(world)]
dist.all_gather(allv, a.value)
assert all(torch.equal(v, allv[0]) for v in allv), "every rank must hold the same iterate"
y_ls = np.linalg.lstsq(w["J"].astype(np.float64), w["d"].astype(np.float64), rcond=None)[0]
scaled = np.abs(a.value.cpu().numpy() - y_ls) / (1e-4 + 1e-4 * np.abs(y_ls))
assert scaled.max() <= 1.0, scaled.max()
if rank == 0:
    print(f"OK world={world} steps={a.stats['num_steps']} max scaled diff vs fp64 lstsq {scaled.max():.3g} loss {a.stats['loss']:.6g}")
dist.barrier()
dist.destroy_process_group()
