#!/usr/bin/env python
"""bench.py - headline benchmark of the B200 Optimistix hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

Headline workload (BASELINE.json configs[1]): LevenbergMarquardt over 1,048,576 independent two-Gaussian +
baseline fits (n=8 parameters, m=256 residuals, fp32, rtol=atol=1e-4, max_steps=256) PER GPU (weak scaling:
problems are independent, ranks share nothing, there is no data-path collective).  A "step" is one pass of
the hot path over that batch.  Prints ONE JSON line (rank 0).

  value     problems/s, whole job, inputs resident in HBM, timed with CUDA events on the launching stream
  e2e       problems/s through the C ABI host entry point (b200opt_least_squares_host) with pinned HOST
            buffers: H2D of y0+data and D2H of y/num_steps/result are inside the timed region
  roofline  the LM kernel against the measured HBM copy bandwidth (MEASURED_PEAKS.json), algorithmic bytes
            = 1096 B/problem; the binding roof is the fp32 FMA pipe, reported in "fp"
  cpu_baseline / --impl reference: the reference's algorithm (augmented QR solve) restated in C + OpenMP
            (oracle/c), timed on this box's host cores on a bounded sample of the same inputs.  The real
            reference (JAX) cannot run in this image.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

WORKLOADS = {
    # name: (description, per-GPU batch, algorithmic HBM bytes / problem, dtype)
    "gauss2": ("LM 2-Gaussian+baseline fits n=8 m=256 fp32 rtol=atol=1e-4 max_steps=256 (BASELINE configs[1])", 1 << 20, 1096, "f32"),
    "exp_decay": ("LM exponential-decay fits n=3 m=64 fp64 rtol=atol=1e-8 max_steps=256 (BASELINE configs[0])", 4096, 568, "f64"),
    "newton": ("Newton implicit-Euler root finds n=32 fp64 rtol=atol=1e-8 (BASELINE configs[2])", 1 << 18, 528, "f64"),
    "bfgs": ("BFGS chained-Rosenbrock minimisations n=64 fp64 rtol=atol=1e-8 max_steps=4096 (BASELINE configs[3])", 1 << 16, 1032, "f64"),
}
CPU_SAMPLE = {"gauss2": 32768, "exp_decay": 4096, "newton": 262144, "bfgs": 16384}


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            p = json.load(fh)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def _traffic(key: str, units: int):
    """DRAM bytes per launch from the committed ncu capture (profiles/traffic.json), scaled per unit to this launch."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
            t = json.load(fh)[key]
        per_unit = t["dram_bytes"] / t["units"]
        return per_unit * units, (f"ncu --set full capture profiles/{t['report']}.csv: {t['dram_bytes'] / 1e6:.1f} MB for {t['units']} {t['unit']}s "
                                  f"= {per_unit:.0f} B/{t['unit']} (algorithmic {t['algorithmic_bytes_per_unit']}), scaled to this launch's {units} {t['unit']}s")
    except Exception:  # noqa: BLE001
        return None, "no ncu capture for this kernel in profiles/traffic.json"


class ClockSampler:
    """Samples SM clock and throttle reasons through NVML during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception as e:  # noqa: BLE001
            self.nv = None
            self.err = repr(e)

    def _run(self):
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:  # noqa: BLE001
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:  # noqa: BLE001
                pass
            self._stop.wait(0.05)

    def start(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()

    def stop(self):
        self._stop.set()
        if self._thr:
            self._thr.join()
        if self.nv is None or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "note": "NVML unavailable"}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------ workloads
def make_inputs(name: str, batch: int, seed_offset: int):
    from optimistix_b200 import synthetic

    if name == "gauss2":
        w = synthetic.gauss2_batch(batch, seed=1 + 1000 * seed_offset)
        return dict(y0=w["y0"], shared=w["x"], prob=w["d"], truth=w["truth"])
    if name == "exp_decay":
        w = synthetic.exp_decay_batch(batch, seed=0 + 1000 * seed_offset)
        return dict(y0=w["y0"], shared=w["t"], prob=w["d"], truth=w["truth"])
    if name == "newton":
        w = synthetic.implicit_euler_batch(batch, seed=2 + 1000 * seed_offset)
        return dict(y0=w["y0"], A=w["A"], B=w["B"], y_prev=w["y_prev"], dt=w["dt"])
    if name == "bfgs":
        w = synthetic.rosenbrock_batch(batch, seed=3 + 1000 * seed_offset)
        return dict(y0=w["y0"])
    raise ValueError(name)


def solver_for(name: str):
    import optimistix_b200 as optx

    if name == "gauss2":
        return optx.models.gauss2, optx.LevenbergMarquardt(1e-4, 1e-4), 256
    if name == "exp_decay":
        return optx.models.exp_decay, optx.LevenbergMarquardt(1e-8, 1e-8), 256
    if name == "newton":
        return optx.models.implicit_euler, optx.Newton(1e-8, 1e-8), 256
    if name == "bfgs":
        return optx.models.rosenbrock, optx.BFGS(1e-8, 1e-8), 4096
    raise ValueError(name)


def run_device(name, inp_dev):
    """One step with inputs resident in HBM: exactly one persistent-kernel launch."""
    import optimistix_b200 as optx

    fn, solver, max_steps = solver_for(name)
    if name in ("gauss2", "exp_decay"):
        return optx.least_squares(fn, solver, inp_dev["y0"], (inp_dev["shared"], inp_dev["prob"]), max_steps=max_steps, throw=False)
    if name == "newton":
        return optx.root_find(fn, solver, inp_dev["y0"], (inp_dev["A"], inp_dev["B"], inp_dev["y_prev"], inp_dev["dt"]),
                              max_steps=max_steps, throw=False)
    return optx.minimise(fn, solver, inp_dev["y0"], None, max_steps=max_steps, throw=False)


def run_host(name, inp_host, chunk):
    """One end-to-end step through the host C entry point: pinned host buffers in, host buffers out."""
    import optimistix_b200 as optx

    fn, solver, max_steps = solver_for(name)
    if name in ("gauss2", "exp_decay"):
        return optx.least_squares(fn, solver, inp_host["y0"], (inp_host["shared"], inp_host["prob"]), max_steps=max_steps,
                                  throw=False, chunk=chunk)
    if name == "newton":
        return optx.root_find(fn, solver, inp_host["y0"], (inp_host["A"], inp_host["B"], inp_host["y_prev"], inp_host["dt"]),
                              max_steps=max_steps, throw=False, chunk=chunk)
    return optx.minimise(fn, solver, inp_host["y0"], None, max_steps=max_steps, throw=False, chunk=chunk)


def bench_gram(args, rank, world, local_rank):
    """BASELINE configs[4]: one large GaussNewton problem, m rows x 256 parameters (fp32), row-sharded over the ranks
    (strong scaling: the total row count is fixed).  A step is one Gauss-Newton iteration:
    fused residual + tensor-core Gram pass, one all-gather of n^2+n+2 floats per rank, Cholesky solve."""
    import torch
    import torch.distributed as dist

    from optimistix_b200._gn_large import N, OUT_FLOATS, CudaOps, PeerExchange

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    m_total = args.rows
    m_local = m_total // world // 64 * 64
    gen = torch.Generator(device=dev)
    gen.manual_seed(4 + rank)
    scale = torch.logspace(0, -2, N, device=dev)
    J = torch.empty((m_local, N), dtype=torch.float32, device=dev)
    chunk = 1 << 20
    for s0 in range(0, m_local, chunk):
        e0 = min(m_local, s0 + chunk)
        J[s0:e0] = torch.randn((e0 - s0, N), generator=gen, device=dev) * scale
    y_true = torch.randn(N, generator=torch.Generator(device=dev).manual_seed(4), device=dev) / scale
    d = torch.empty(m_local, dtype=torch.float32, device=dev)
    for s0 in range(0, m_local, chunk):
        e0 = min(m_local, s0 + chunk)
        d[s0:e0] = J[s0:e0] @ y_true + 0.01 * torch.randn(e0 - s0, generator=gen, device=dev)
    ops = CudaOps(dev)
    y = torch.zeros(N, device=dev)
    r = torch.empty(m_local, device=dev)
    r_old = torch.zeros(m_local, device=dev)
    out = torch.empty(OUT_FLOATS, device=dev)  # this rank's [G | g | rr | f-stat]
    fmax = out[N * N + N + 1:].view(torch.int32)
    gathered = torch.empty(world * OUT_FLOATS, device=dev) if world > 1 else out
    summary = torch.zeros(2, device=dev)
    p = torch.empty(N, device=dev)
    status = torch.zeros(1, dtype=torch.int32, device=dev)

    fused = not args.unfused
    xchg = None
    if world > 1 and fused and args.exchange == "peer":
        xchg = PeerExchange(dev)  # CUDA IPC mapping of every rank's exchange buffer; handles swapped once through torch.distributed

    def step():
        if xchg is not None:
            parity, seq = xchg.next()
            ops.residual_gram_xchg(J, y, d, r, r_old, 1e-4, 1e-4, fmax, out, xchg, parity, seq)  # Gram + publication over NVLink
            ops.solve_xchg(xchg, parity, seq, 0.0, p, status, summary)                            # wait for all flags, sum, solve
            return
        if fused:
            ops.residual_gram(J, y, d, r, r_old, 1e-4, 1e-4, fmax, out)
        else:
            ops.residual(J, y, d, r, r_old, 1e-4, 1e-4, fmax)
            ops.gram(J, r, out)
        if world > 1:
            dist.all_gather_into_tensor(gathered, out)  # the only exchange: n^2 + n + 2 floats per rank
        ops.solve_parts(gathered, world, OUT_FLOATS, 0.0, p, status, summary)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    W, K = max(args.warmup, 3), max(args.steps, 1)
    for _ in range(W):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    # K steps enqueued back to back (the host runs ahead; nothing synchronises inside the timed region), 4 events per step
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(K)]
    for ev in evs:
        ev[0].record()
        if not fused:
            ops.residual(J, y, d, r, r_old, 1e-4, 1e-4, fmax)
        ev[1].record()
        if xchg is not None:
            parity, seq = xchg.next()
            ops.residual_gram_xchg(J, y, d, r, r_old, 1e-4, 1e-4, fmax, out, xchg, parity, seq)
        elif fused:
            ops.residual_gram(J, y, d, r, r_old, 1e-4, 1e-4, fmax, out)
        else:
            ops.gram(J, r, out)
        ev[2].record()
        if xchg is not None:
            ops.solve_xchg(xchg, parity, seq, 0.0, p, status, summary)
        else:
            if world > 1:
                dist.all_gather_into_tensor(gathered, out)
            ops.solve_parts(gathered, world, OUT_FLOATS, 0.0, p, status, summary)
        ev[3].record()
    torch.cuda.synchronize()
    t_step = evs[0][0].elapsed_time(evs[-1][3])  # first launch to last completion: the K steps including any gaps
    t_res = sum(ev[0].elapsed_time(ev[1]) for ev in evs)
    t_gram = sum(ev[1].elapsed_time(ev[2]) for ev in evs)
    barrier()
    clocks = sampler.stop()
    # size-independent property at the full size (outside the timed region): Gauss-Newton from y = 0 must recover the
    # parameters the data were generated from (d = J y_true + 0.01 N(0,1)) to the noise level, on every rank identically
    y.zero_()
    r_old.zero_()
    gn_steps = 0
    for gn_steps in range(1, 13):
        step()
        y_conv = bool(((p.abs() / (1e-4 + 1e-4 * y.abs())).max() < 1).item()) if gn_steps > 1 else False  # Cauchy test on y ...
        f_conv = float(summary[1].item()) < 1.0  # ... and on the residuals (_misc.py:291-310), as gauss_newton_large does
        if y_conv and f_conv:
            break
        y.sub_(p)
        r, r_old = r_old, r
    # the least-squares estimate scatters around y_true with std sigma / (sqrt(m) * column scale): report the largest
    # z-score over the 256 parameters (about 3 for a correct solve; a 1e-4 relative numerical error would read >> 10)
    rel = ((y - y_true).abs() * scale * (float(m_local * world) ** 0.5) / 0.01).max()
    if world > 1:
        dist.all_reduce(rel, op=dist.ReduceOp.MAX)
    checks = {"gn_steps_to_convergence": gn_steps, "max_zscore_vs_generating_parameters": float(rel.item()),
              "loss": 0.5 * float(summary[0].item()), "linear_solve_status": int(status.item())}
    if world > 1:
        t = torch.tensor([t_step, t_gram, t_res], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_step, t_gram, t_res = (float(x) for x in t)
    ms, ms_gram, ms_res = t_step / K, t_gram / K, t_res / K
    if rank == 0:
        rows = m_local * world
        peak, peak_src = _peaks()
        # algorithmic bytes per row: J 1024 (+ d, r_old read and r written when the residual is fused; + r read when not)
        alg_bytes = m_local * (N * 4 + (12 if fused else 4))
        traffic, traffic_note = _traffic("gram_kernel(fused)", m_local) if fused else (None, "no capture of the unfused variant this round")
        tf32_peak = None
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
                tf32_peak = float(json.load(fh)["bf16_tflops"]) / 2
        except Exception:  # noqa: BLE001
            pass
        gram_tflops = 2.0 * m_local * N * N / (ms_gram / 1e3) / 1e12  # per GPU, full (unsymmetrised) count
        line = {"metric": "GaussNewton Gram iterations/sec (m rows x 256 params, row-sharded)", "value": 1e3 / ms, "unit": "iterations/s",
                "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "tf32 x tf32 -> f32 (TMEM)", "data": "synthetic",
                "config": {"workload": "GaussNewton least squares, J materialised row-major in HBM (BASELINE configs[4])", "rows_total": rows,
                           "fused_residual": fused,
                           "rows_per_gpu": m_local, "n": N, "parallelism": (f"row-sharded x{world}; per iteration each rank's {N * N + N + 2}-float block is "
                                           + ("stored into every peer's exchange buffer by the Gram reduction kernel (P2P over NVLink, flag-signalled)"
                                              if xchg is not None else "all-gathered (NCCL)")
                                           + ", summed in rank order inside the solve kernel"),
                           "l2": "J shard (>= 8 GB) exceeds the 126 MB L2"},
                "roofline": {"bound": "hbm", "achieved": alg_bytes / (ms_gram / 1e3) / 1e9, "peak": peak, "unit": "GB/s",
                             "frac": alg_bytes / (ms_gram / 1e3) / 1e9 / peak, "traffic": traffic, "traffic_note": traffic_note,
                             "kernel": "gram_kernel", "peak_source": peak_src, "algorithmic_bytes_per_row": alg_bytes // m_local},
                "tensor": {"achieved_tflops_full_2mn2": gram_tflops, "executed_fraction_of_full": 0.75,
                           "peak_tflops_tf32_assumed": tf32_peak, "frac": (gram_tflops / tf32_peak) if tf32_peak else None,
                           "note": "tf32 peak assumed = measured bf16 peak / 2 (no tf32 measurement in MEASURED_PEAKS.json)"},
                "breakdown_ms": {"residual_pass": ms_res, "gram": ms_gram, "allgather+solve": ms - ms_res - ms_gram},
                "gpu_launches": (3 if fused else 4) * K, "clocks": clocks, "checks": checks, "cpu_baseline": None, "e2e": None}
        print(json.dumps(line))
    if xchg is not None:
        barrier()
        xchg.close()
    if world > 1:
        dist.destroy_process_group()


def cpu_reference(name, inp, sample, linsolve="qr"):
    """The reference algorithm restated on the CPU (oracle), all host threads, on the first `sample` problems."""
    from oracle import c_oracle

    if name in ("gauss2", "exp_decay"):
        tol = 1e-4 if name == "gauss2" else 1e-8
        # all host threads this process may run on (torchrun exports OMP_NUM_THREADS=1, so ask the OS, not OpenMP)
        try:
            cores = len(os.sched_getaffinity(0))
        except AttributeError:
            cores = os.cpu_count() or 1
        t0 = time.perf_counter()
        out = c_oracle.levenberg_marquardt(name, inp["y0"][:sample], inp["shared"], inp["prob"][:sample], rtol=tol, atol=tol,
                                           max_steps=256, linsolve=linsolve, nthreads=cores)
        dt = time.perf_counter() - t0
        return out, dt, cores, "C+OpenMP restatement (oracle/c), augmented-QR solve as in the reference"
    try:
        cores = len(os.sched_getaffinity(0))
    except AttributeError:
        cores = os.cpu_count() or 1
    t0 = time.perf_counter()
    if name == "newton":
        out = c_oracle.newton_implicit_euler(inp["y0"][:sample], inp["A"], inp["B"], inp["y_prev"][:sample], inp["dt"][:sample],
                                             rtol=1e-8, atol=1e-8, nthreads=cores)
    else:
        out = c_oracle.bfgs_rosenbrock(inp["y0"][:sample], rtol=1e-8, atol=1e-8, max_steps=4096, nthreads=cores)
    dt = time.perf_counter() - t0
    return out, dt, cores, "C+OpenMP restatement (oracle/c/nb_oracle.c) of the reference step, cross-checked against the NumPy oracle"


# ------------------------------------------------------------------------------------------------ main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="gauss2", choices=sorted(WORKLOADS) + ["gram"])
    ap.add_argument("--rows", type=int, default=64 << 20, help="gram workload: total rows of J over all ranks")
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch override (default: the BASELINE config)")
    ap.add_argument("--chunk", type=int, default=1 << 15, help="problems per pipeline stage in the host entry point")
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"],
                    help="gram workload, N > 1: peer = our kernels exchange the Gram blocks over peer memory; nccl = all_gather_into_tensor")
    ap.add_argument("--unfused", action="store_true", help="gram workload: separate residual pass (J read twice per iteration)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    name = args.workload
    if name == "gram":
        if args.impl == "reference":
            if rank == 0:
                print(json.dumps({"impl": "reference", "unavailable": "gram workload has no CPU reference arm (68.7 GB QR); see tests/test_gram_gpu.py for parity"}))
            return
        return bench_gram(args, rank, world, local_rank)
    desc, batch, bytes_per_problem, dtype = WORKLOADS[name]
    if args.batch:
        batch = args.batch
    W = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    K = max(args.steps, 1)
    config = {"workload": desc, "batch_per_gpu": batch, "global_batch": batch * world, "parallelism": f"batch-sharded x{world}, no collective",
              "l2": "inputs (>= 1 GB at the headline size) exceed the 126 MB L2; every step re-reads them from HBM"}

    # ---------------------------------------------------------------- reference arm: CPU only, rank 0 only
    if args.impl == "reference":
        if rank != 0:
            return
        sample = min(CPU_SAMPLE[name], batch)
        inp = make_inputs(name, sample, 0)
        for _ in range(args.warmup):
            cpu_reference(name, inp, sample)
        t = []
        for _ in range(K):
            _, dt, cores, kind_note = cpu_reference(name, inp, sample)
            t.append(dt)
        ms = 1e3 * float(np.mean(t))
        val = sample / (ms / 1e3)
        line = {"impl": "reference", "metric": "LM problems solved/sec" if name in ("gauss2", "exp_decay") else f"{name} problems solved/sec",
                "value": val, "unit": "problems/s", "n_gpus": args.gpus, "steps": K, "warmup": args.warmup, "ms_per_step": ms,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": dtype, "data": "synthetic",
                "config": config,
                "cpu_baseline": {"value": val, "unit": "problems/s", "cores": cores, "kind": "port",
                                 "sample": f"first {sample} problems of the workload per step; {kind_note}; the JAX reference itself cannot run in this image"},
                "e2e": {"value": val, "unit": "problems/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    # ---------------------------------------------------------------- our arm
    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    inp = make_inputs(name, batch, rank)
    keys = [k for k in inp if k != "truth"]
    inp_dev = {k: torch.as_tensor(inp[k], device=dev) for k in keys}
    torch.cuda.synchronize()

    # ---- value: device-resident, one kernel launch per step, CUDA events on the launching (current) stream
    for _ in range(W):
        sol = run_device(name, inp_dev)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    t_wall0 = time.perf_counter()
    for a, b in evs:
        a.record()
        sol = run_device(name, inp_dev)
        b.record()
    torch.cuda.synchronize()
    t_total_local = sum(a.elapsed_time(b) for a, b in evs)  # ms, device time of the K steps
    barrier()
    wall_ms = 1e3 * (time.perf_counter() - t_wall0)
    clocks = sampler.stop()
    t_total = max_over_ranks(t_total_local)
    ms_per_step = t_total / K
    value = batch * world / (ms_per_step / 1e3)
    launches = K

    # size-independent checks at the full size (cheap, outside the timed region)
    res = sol.result
    steps_t = sol.stats["num_steps"]
    n_ok = int((res == 0).sum().item())
    mean_steps = float(steps_t.float().mean().item())
    checks = {"successful_frac": n_ok / batch, "mean_num_steps": mean_steps}
    if "truth" in inp:
        err = (sol.value - torch.as_tensor(inp["truth"], device=dev)).abs().median(0).values
        checks["median_abs_err_vs_truth_max"] = float(err.max().item())

    # ---- e2e: host C entry point, pinned host buffers, copies inside the timed region
    e2e = None
    if not args.no_e2e:
        inp_host = {k: torch.as_tensor(inp[k]).pin_memory().numpy() for k in keys}
        h2d = sum(inp_host[k].nbytes for k in keys)
        for _ in range(2):
            hs = run_host(name, inp_host, args.chunk)
        barrier()
        t0 = time.perf_counter()
        for _ in range(K):
            hs = run_host(name, inp_host, args.chunk)
        torch.cuda.synchronize()
        e2e_ms = max_over_ranks(1e3 * (time.perf_counter() - t0)) / K
        d2h = hs.value.nbytes + hs.stats["num_steps"].nbytes + hs.result.nbytes
        same = bool(np.array_equal(hs.value, sol.value.cpu().numpy()))
        checks["e2e_bitwise_equal_device_path"] = same
        e2e = {"value": batch * world / (e2e_ms / 1e3), "unit": "problems/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "ms_per_step": e2e_ms, "api": "b200opt_least_squares_host via optimistix_b200.least_squares(numpy inputs)" if name in ("gauss2", "exp_decay") else "host entry point",
               "chunk": args.chunk}
        launches += K * ((batch + args.chunk - 1) // args.chunk)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant (only) kernel
    peak, peak_src = _peaks()
    achieved = bytes_per_problem * batch / (ms_per_step / 1e3) / 1e9  # GB/s per GPU (per launch)
    tkey = {"gauss2": "lsq_solve_kernel<Gauss2<float>>", "newton": "root_solve_kernel<ImplicitEuler<double>>",
            "bfgs": "bfgs_solve_kernel<Rosenbrock<double>>"}.get(name, "")
    traffic, traffic_note = _traffic(tkey, batch)
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "traffic_note": traffic_note,
                "kernel": "lsq_solve_kernel" if name in ("gauss2", "exp_decay") else name,
                "algorithmic_bytes_per_problem": bytes_per_problem, "peak_source": peak_src,
                "note": "HBM is NOT the binding roof for this path (SURVEY F8): state is on-chip, ~0.5 Mflop per 1.1 KB; see 'fp'"}
    flops_per_iter = {"gauss2": 50e3, "exp_decay": 6.5e3, "newton": 28e3, "bfgs": 41e3}[name]
    fp_achieved = value / world * mean_steps * flops_per_iter / 1e12
    fp_peak = 148 * 128 * 2 * 1.965e9 / 1e12 * (1.0 if dtype == "f32" else 0.5)
    fp = {"achieved_tflops_algorithmic": fp_achieved, "peak_tflops_nominal": fp_peak, "frac": fp_achieved / fp_peak,
          "note": "algorithmic flops/iteration from SURVEY 8d x measured mean iterations; nominal FMA-pipe peak at max clock"}

    cpu = None
    if not args.no_cpu and world == 1:
        sample = min(CPU_SAMPLE[name], batch)
        out, dt, cores, kind_note = cpu_reference(name, inp, sample)
        cpu = {"value": sample / dt, "unit": "problems/s", "cores": cores, "kind": "port",
               "sample": f"first {sample} problems of the same inputs, one pass; {kind_note}"}
        gv = sol.value[:sample].cpu().numpy()
        tol = 1e-4 if dtype == "f32" else 1e-8
        scaled = np.max(np.abs(gv - out["value"]) / (tol + tol * np.abs(out["value"])), axis=1)
        checks["gpu_vs_cpu_port_within_tol_frac"] = float(np.mean(scaled <= 1.0))
        checks["cpu_port_mean_num_steps"] = float(out["num_steps"].mean())

    line = {"metric": "LM problems solved/sec" if name in ("gauss2", "exp_decay") else f"{name} problems solved/sec",
            "value": value, "unit": "problems/s", "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": dtype, "data": "synthetic",
            "config": config, "roofline": roofline, "fp": fp, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches,
            "clocks": clocks, "checks": checks, "wall_ms_timed_region": wall_ms}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
