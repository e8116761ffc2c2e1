"""Multi-rank host logic of the row-sharded Gauss-Newton path on CPU: world_size 2, gloo, 127.0.0.1.

The CUDA kernels are replaced by NumPy stand-ins (test infrastructure, same contract as CudaOps) so that what is
tested is exactly what cannot be tested on one GPU: row sharding, the all-reduce of [G | g | rr] and of the f-test
statistic, and that every rank walks the same iteration sequence to the same answer as the unsharded solve."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import optimistix_b200 as optx
from optimistix_b200 import synthetic
from optimistix_b200._gn_large import N, gauss_newton_large


class NumpyOps:
    launches = 0

    def residual(self, J, y, d, r_out, r_old, rtol, atol, fmax):
        r = J @ y - d
        stat = ((r - r_old).abs() / (atol + rtol * r.abs())).max() if r_old is not None else torch.zeros(())
        r_out.copy_(r)
        fmax.view(torch.float32)[0] = stat  # `fmax` is a view of the last slot of this rank's block

    def gram(self, J, r, out):
        Jd = J.double()
        out[: N * N] = (Jd.T @ Jd).float().reshape(-1)
        out[N * N : N * N + N] = (Jd.T @ r.double()).float()
        out[N * N + N] = float(r.double() @ r.double())

    def solve_parts(self, parts, nparts, stride, lam, p, status, summary):
        blocks = parts.view(nparts, stride).double()
        tot = blocks.sum(0)  # rank order, fp64 (what b200opt_gn_solve_parts does)
        G = tot[: N * N].view(N, N) + lam * torch.eye(N, dtype=torch.float64)
        x = torch.linalg.solve(G, tot[N * N : N * N + N])
        p.copy_(x.float())
        status[0] = 0 if bool(torch.isfinite(x).all()) else 2
        summary[0] = float(tot[N * N + N])
        summary[1] = parts.view(nparts, stride)[:, N * N + N + 1].max()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, m, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    w = synthetic.linear_lsq_rows(m, seed=4)
    lo, hi = rank * m // world, (rank + 1) * m // world
    J, d = torch.as_tensor(w["J"][lo:hi]), torch.as_tensor(w["d"][lo:hi])
    sol = gauss_newton_large(J, d, torch.zeros(N), optx.GaussNewton(1e-4, 1e-4), throw=False, ops=NumpyOps())
    q.put((rank, sol.value.numpy(), sol.stats["num_steps"], sol.result, sol.stats["loss"]))
    dist.destroy_process_group()


def test_row_sharded_gauss_newton_two_ranks():
    m = 4096
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, m, q)) for r in range(2)]
    [p.start() for p in procs]
    outs = sorted(q.get(timeout=120) for _ in procs)
    [p.join(timeout=60) for p in procs]
    w = synthetic.linear_lsq_rows(m, seed=4)
    single = gauss_newton_large(torch.as_tensor(w["J"]), torch.as_tensor(w["d"]), torch.zeros(N), optx.GaussNewton(1e-4, 1e-4),
                                throw=False, ops=NumpyOps())
    y_ls = np.linalg.lstsq(w["J"].astype(np.float64), w["d"].astype(np.float64), rcond=None)[0]
    for rank, y, steps, result, loss in outs:
        assert result == optx.RESULTS.successful
        assert steps == single.stats["num_steps"]
        assert np.allclose(y, single.value.numpy(), rtol=1e-5, atol=1e-6)
        assert np.allclose(y, y_ls, rtol=1e-3, atol=1e-4)
        assert abs(loss - single.stats["loss"]) <= 1e-4 * abs(single.stats["loss"])
    assert np.array_equal(outs[0][1], outs[1][1])  # every rank holds the same iterate
