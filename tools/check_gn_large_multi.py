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
a = optx.least_squares(optx.models.linear_rows, optx.GaussNewton(1e-4, 1e-4), y0, (J, d), throw=False)  # default: peer exchange
b = gauss_newton_large(J, d, y0, optx.GaussNewton(1e-4, 1e-4), throw=False, exchange="nccl")
assert a.result == b.result == optx.RESULTS.successful, (a.result, b.result)
assert a.stats["num_steps"] == b.stats["num_steps"]
assert torch.equal(a.value, b.value), "peer exchange and NCCL all-gather must give the same bits"
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
