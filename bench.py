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
  secondary the other BASELINE configs measured in the SAME run (so that every number of DESIGN.md is in a driver
            file): gauss2_strong (configs[1] as BASELINE words it: 1 Mi fits in total, batch-sharded over the N
            ranks), exp_decay (configs[0]), newton (configs[2]), bfgs (configs[3]) - each strong-scaled over the N
            ranks with value / e2e / roofline / parity against the C port - and gram (configs[4]: 64 Mi rows
            row-sharded over the N ranks, timed THROUGH optx.least_squares, peer-memory and NCCL exchange,
            breakdown_ms, HBM fraction, tensor fraction on executed flops against a TF32 GEMM peak measured in
            the same run).  `--no-secondary` skips it; `--workload X` runs one workload as the headline.
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
        self._armed = threading.Event()  # samples are kept only after arm(): the thread itself starts before the warm-up
        self._thr = None
        try:
            if os.environ.get("BENCH_NO_SAMPLER"):
                raise RuntimeError("disabled by BENCH_NO_SAMPLER")
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
                mhz = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:  # noqa: BLE001
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                if self._armed.is_set():
                    self.samples.append(mhz)
                    for bit, name in names.items():
                        if mask & bit:
                            self.reasons.add(name)
            except Exception:  # noqa: BLE001
                pass
            self._stop.wait(0.02)

    def start(self, armed: bool = True):
        """armed=False: start the thread now (before the warm-up: the first NVML queries of a process take tens to hundreds of
        milliseconds and serialise with CUDA launches - measured as a 0.5 s hole in a timed region), keep samples from arm() on."""
        if armed:
            self._armed.set()
        if self.nv is not None:
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()

    def arm(self):
        self._armed.set()

    def stop(self):
        self._stop.set()
        if self._thr:
            self._thr.join()
        note = None
        if self.nv is not None and not self.samples:
            # the timed region was shorter than the sampling period: one query right at its end (the SM clock of a GPU that was
            # busy microseconds ago; the reasons mask is the current one)
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                note = "timed region shorter than the 20 ms sampling period: sampled once at its end"
            except Exception:  # noqa: BLE001
                pass
        if self.nv is None or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "note": "NVML unavailable"}
        if note:
            return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                    "samples": len(self.samples), "note": note}
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


def _tf32_gemm_peak(dev):
    """cuBLAS TF32 GEMM throughput measured in this run (8192^3, best of 10): the denominator of the tensor fraction."""
    import torch

    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        n = 8192
        a = torch.randn(n, n, device=dev)
        b = torch.randn(n, n, device=dev)
        c = torch.empty(n, n, device=dev)
        for _ in range(3):
            torch.matmul(a, b, out=c)
        best = float("inf")
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            torch.matmul(a, b, out=c)
            e1.record()
            e1.synchronize()
            best = min(best, e0.elapsed_time(e1))
        return 2.0 * n**3 / (best / 1e3) / 1e12
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old


def bench_gram(args, rank, world, dev, K, W):
    """BASELINE configs[4]: one large GaussNewton problem, m rows x 256 parameters (fp32), row-sharded over the ranks
    (strong scaling: the total row count is fixed).  A step is one Gauss-Newton iteration.  Timed THROUGH the public API:
    optx.least_squares(models.linear_rows, GaussNewton(rtol=atol=0), ..., max_steps=K) runs exactly K iterations (a zero
    tolerance never terminates), with everything an iteration does (residual, Gram, exchange, solve, y update, Cauchy test,
    step counter) inside the timed region.  `kernel_only` repeats the round-1 measurement (Gram + separate reduce kernel + fp64
    solve enqueued back to back through the step-wise C entry points, nothing else) for the same exchange so that the two can
    be compared; `solve_kernel_ms` times the two 256 x 256 solvers alone."""
    import torch
    import torch.distributed as dist

    import optimistix_b200 as optx
    from optimistix_b200._gn_large import N, OUT_FLOATS, CudaOps, PeerExchange

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
    y0 = torch.zeros(N, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(vals):
        t = torch.tensor(vals, dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(x) for x in t]

    never = optx.GaussNewton(rtol=0.0, atol=0.0)  # |diff| / 0 < 1 is never true: exactly max_steps iterations

    def api_run(steps, exchange, linear_solver, device_loop=True):
        from optimistix_b200._gn_large import gauss_newton_large

        return gauss_newton_large(J, d, y0, never, max_steps=steps, throw=False, exchange=exchange, linear_solver=linear_solver,
                                  device_loop=device_loop, check_every=max(steps, 1))

    def time_api(exchange, linear_solver, device_loop=True):
        api_run(W, exchange, linear_solver, device_loop)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        sol = api_run(K, exchange, linear_solver, device_loop)
        e1.record()
        torch.cuda.synchronize()
        wall = 1e3 * (time.perf_counter() - t0)
        dev_ms = e0.elapsed_time(e1)
        barrier()
        assert sol.stats["num_steps"] == K, sol.stats
        dev_ms, wall = max_over_ranks([dev_ms, wall])
        return dev_ms / K, wall / K, sol

    # ---- kernel-only (round-1 style): K x [Gram (+ reduce) | solve] enqueued back to back through the step-wise C entry points
    ops = CudaOps(dev)
    y = torch.zeros(N, device=dev)
    r = torch.empty(m_local, device=dev)
    r_old = torch.zeros(m_local, device=dev)
    out = torch.empty(OUT_FLOATS, device=dev)
    fmax = out[N * N + N + 1:].view(torch.int32)
    gathered = torch.empty(world * OUT_FLOATS, device=dev) if world > 1 else out
    summary = torch.zeros(2, device=dev)
    p = torch.empty(N, device=dev)
    status = torch.zeros(1, dtype=torch.int32, device=dev)

    def time_kernels(exchange):
        xchg = PeerExchange(dev) if (world > 1 and exchange == "peer") else None

        def step(evs=None):
            if evs:
                evs[0].record()
            if xchg is not None:
                parity, seq = xchg.next()
                ops.residual_gram_xchg(J, y, d, r, r_old, 1e-4, 1e-4, fmax, out, xchg, parity, seq)
            else:
                ops.residual_gram(J, y, d, r, r_old, 1e-4, 1e-4, fmax, out)
            if evs:
                evs[1].record()
            if xchg is not None:
                ops.solve_xchg(xchg, parity, seq, 0.0, p, status, summary)
            else:
                if world > 1:
                    dist.all_gather_into_tensor(gathered, out)
                ops.solve_parts(gathered, world, OUT_FLOATS, 0.0, p, status, summary)
            if evs:
                evs[2].record()

        for _ in range(W):
            step()
        barrier()
        evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(K)]
        for ev in evs:
            step(ev)
        torch.cuda.synchronize()
        t_step = evs[0][0].elapsed_time(evs[-1][2])
        t_gram = sum(ev[0].elapsed_time(ev[1]) for ev in evs)
        barrier()
        if xchg is not None:
            xchg.close()
        t_step, t_gram = max_over_ranks([t_step, t_gram])
        return t_step / K, t_gram / K

    def time_solvers():
        """The 256 x 256 solve alone (b200opt_gn_solve_fast on `world` copies of this rank's block): fp32 + refinement vs fp64."""
        lib = ops.lib
        stride = OUT_FLOATS + 2  # a multiple of 4
        parts = torch.zeros(world * stride, device=dev)
        for q in range(world):
            parts[q * stride: q * stride + OUT_FLOATS] = out
        info = torch.zeros(1, dtype=torch.int32, device=dev)
        tms = {}
        for nm, sid in (("fp32_blocked_cholesky+fp64_refinement", 0), ("fp64_cholesky", 1)):
            def one():
                return lib.b200opt_gn_solve_fast(parts.data_ptr(), world, stride, N, 0.0, sid, p.data_ptr(), status.data_ptr(), summary.data_ptr(),
                                                 info.data_ptr(), ops.ws.data_ptr(), torch.cuda.current_stream(dev).cuda_stream)
            for _ in range(3):
                one()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                one()
            e1.record()
            torch.cuda.synchronize()
            tms[nm] = e0.elapsed_time(e1) / 20
        tms["fp32_path_accepted"] = None
        lib.b200opt_gn_solve_fast(parts.data_ptr(), world, stride, N, 0.0, 0, p.data_ptr(), status.data_ptr(), summary.data_ptr(), info.data_ptr(),
                                  ops.ws.data_ptr(), torch.cuda.current_stream(dev).cuda_stream)
        tms["fp32_path_accepted"] = bool(int(info.item()))
        return tms

    sampler = ClockSampler(dev.index)
    sampler.start()
    res = {}
    exchanges = ["peer", "nccl"] if world > 1 else ["local"]
    # every configuration is timed twice, interleaved with the others (A B C A B C), and the faster pass is kept: under
    # a power cap the SM clock drifts over seconds, which would otherwise be read as a difference between configurations
    for rep_i in range(2):
        for ex in exchanges:
            entry = res.setdefault(ex, {})

            def keep(key, val):
                entry[key] = min(entry.get(key, float("inf")), val)

            ms_api, wall_api, sol = time_api(ex if world > 1 else "peer", 0)
            keep("api_ms_per_iteration", ms_api)
            keep("api_wall_ms_per_iteration", wall_api)
            entry["fp64_fallbacks"] = sol.stats.get("fp64_fallbacks")
            entry["api_path"] = sol.stats.get("exchange")
            if ex != "nccl":
                keep("api_ms_per_iteration_fp64_solver", time_api(ex if world > 1 else "peer", 1)[0])
            ms_k, ms_g = time_kernels(ex)
            if ms_k < entry.get("kernel_only_ms_per_iteration", float("inf")):
                entry["kernel_only_ms_per_iteration"] = ms_k
                entry["kernel_only_breakdown_ms"] = {"gram+reduce(+publish)": ms_g, "exchange_wait+solve": ms_k - ms_g}
    for entry in res.values():
        entry["api_over_kernel_only"] = entry["api_ms_per_iteration"] / entry["kernel_only_ms_per_iteration"]
    solve_ms = time_solvers()
    clocks = sampler.stop()
    best = min(res, key=lambda k: res[k]["api_ms_per_iteration"])
    ms = res[best]["api_ms_per_iteration"]

    # size-independent property at the full size (outside the timed region): Gauss-Newton from y = 0 through the public API
    # must converge and recover the generating parameters to the noise level, bit-identically on every rank
    sol = optx.least_squares(optx.models.linear_rows, optx.GaussNewton(1e-4, 1e-4), y0, (J, d), throw=False)
    rel = ((sol.value - y_true).abs() * scale * (float(m_local * world) ** 0.5) / 0.01).max()
    same = torch.ones(1, device=dev)
    if world > 1:
        dist.all_reduce(rel, op=dist.ReduceOp.MAX)
        ref = sol.value.clone()
        dist.broadcast(ref, 0)
        same = torch.tensor([float(torch.equal(ref, sol.value))], device=dev)
        dist.all_reduce(same, op=dist.ReduceOp.MIN)
    checks = {"gn_steps_to_convergence": int(sol.stats["num_steps"]), "result": optx.RESULTS.name(int(sol.result)),
              "max_zscore_vs_generating_parameters": float(rel.item()), "loss": float(sol.stats["loss"]),
              "loss_expected_0.5_m_sigma2": 0.5 * m_local * world * 1e-4, "iterate_bit_identical_on_all_ranks": bool(same.item()),
              "fp64_fallbacks": sol.stats.get("fp64_fallbacks")}
    tf32_peak = _tf32_gemm_peak(dev) if rank == 0 else None
    if rank != 0:
        return None
    peak, peak_src = _peaks()
    rows = m_local * world
    # the Gram pass alone (kernel-only breakdown of the best exchange) against the HBM roof; algorithmic bytes per row:
    # J 1024 + d 4 + r_old 4 + r written 4
    ms_gram = res[best]["kernel_only_breakdown_ms"]["gram+reduce(+publish)"]
    alg_bytes = m_local * (N * 4 + 12)
    traffic, traffic_note = _traffic("gram_kernel(fused)", m_local)
    full_tflops = 2.0 * m_local * N * N / (ms_gram / 1e3) / 1e12
    line = {"metric": "GaussNewton iterations/sec (m rows x 256 params, row-sharded; through optx.least_squares)", "value": 1e3 / ms,
            "unit": "iterations/s", "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "tf32 x tf32 -> f32 (TMEM); fp32 blocked Cholesky + fp64 refinement", "data": "synthetic",
            "config": {"workload": "GaussNewton least squares, J materialised row-major in HBM (BASELINE configs[4])", "rows_total": rows,
                       "rows_per_gpu": m_local, "n": N, "exchange_of_value": best,
                       "parallelism": f"row-sharded x{world}; per iteration each rank's {N * N + N + 2}-float block is stored into every peer's exchange "
                                      "buffer by the Gram kernel itself (P2P over NVLink, flag-signalled) and summed in rank order inside the solve kernel"
                                      if world > 1 else "single GPU",
                       "l2": "J shard (>= 8 GB) exceeds the 126 MB L2"},
            "exchanges": res, "solve_kernel_ms": solve_ms,
            "roofline": {"bound": "hbm", "achieved": alg_bytes / (ms_gram / 1e3) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": alg_bytes / (ms_gram / 1e3) / 1e9 / peak, "traffic": traffic, "traffic_note": traffic_note,
                         "kernel": "gram_kernel (fused residual + Gram pass)", "peak_source": peak_src, "algorithmic_bytes_per_row": alg_bytes // m_local,
                         "whole_iteration_frac": alg_bytes / (ms / 1e3) / 1e9 / peak},
            "tensor": {"executed_tflops": 0.75 * full_tflops, "full_2mn2_tflops": full_tflops, "executed_fraction_of_full": 0.75,
                       "peak_tflops_tf32_measured": tf32_peak, "frac": 0.75 * full_tflops / tf32_peak if tf32_peak else None,
                       "note": "executed flops = 0.75 x 2 m n^2 (the symmetric half of one 128 x 128 block is skipped); peak = cuBLAS TF32 8192^3 GEMM measured in this run"},
            "breakdown_ms": {"api_iteration": ms, **res[best]["kernel_only_breakdown_ms"]},
            "gpu_launches": 2 * K, "clocks": clocks, "checks": checks, "cpu_baseline": None,
            "e2e": {"value": 1e3 / res[best]["api_wall_ms_per_iteration"], "unit": "iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 48 / max(K, 1),
                    "note": "host wall clock around the same public-API call (J lives on the device: 68.7 GB; the only host traffic is the 48-byte state read)"}}
    return line


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


# ------------------------------------------------------------------------------------------------ batched workloads
def _fp_peak(dtype):
    """Measured FMA-pipe peak (profiles/microbench.json, tools/microbench.cu on a B200 of this pool): packed fp32 / fp64."""
    try:
        with open(os.path.join(ROOT, "profiles", "microbench.json")) as fh:
            mb = json.load(fh)
        if dtype == "f32":
            return float(mb["ffma2_packed"]["tflops_if_fma"]), "measured (profiles/microbench.json ffma2_packed)"
        return float(mb["dfma"]["tflops_if_fma"]), "measured (profiles/microbench.json dfma)"
    except Exception:  # noqa: BLE001
        return 148 * 128 * 2 * 1.965e9 / 1e12 * (1.0 if dtype == "f32" else 0.5), "nominal (no profiles/microbench.json)"


def bench_batched(name, batch, inp, args, rank, world, dev, K, W, do_e2e, do_cpu):
    """One batched workload on this rank's `batch` problems (inputs `inp`, NumPy, already generated): device-resident value,
    e2e through the host C entry point, roofline, parity checks.  Returns the result dict on rank 0, None elsewhere."""
    import torch
    import torch.distributed as dist

    desc, _, bytes_per_problem, dtype = WORKLOADS[name]

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

    keys = [k for k in inp if k != "truth"]
    inp_dev = {k: torch.as_tensor(inp[k], device=dev) for k in keys}
    torch.cuda.synchronize()

    # ---- value: device-resident, one kernel launch per step, CUDA events on the launching (current) stream
    sampler = ClockSampler(dev.index)
    sampler.start(armed=False)
    for _ in range(W):
        sol = run_device(name, inp_dev)
    # a step of the small configs lasts ~0.1 ms: keep warming up until the GPU has been busy for 30 ms, or the clocks
    # (idle during the host-side input generation) are still ramping inside the timed region
    torch.cuda.synchronize()
    t_warm = time.perf_counter()
    while time.perf_counter() - t_warm < 0.03:
        sol = run_device(name, inp_dev)
        torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    for a, b in evs:  # CUDA events are created lazily by their first record(): do that outside the timed region
        a.record(), b.record()
    # the timed region starts from an idle GPU (contract), so the first step's bracket contains that step's host-side launch
    # path: keep the interpreter's collector out of it (measured: occasional 4 - 20 ms holes in the first bracket)
    import gc

    gc.collect()
    gc.disable()
    torch.cuda.synchronize()
    barrier()
    sampler.arm()
    t_wall0 = time.perf_counter()
    host_us = []
    for a, b in evs:
        th = time.perf_counter()
        a.record()
        sol = run_device(name, inp_dev)
        b.record()
        host_us.append(round(1e6 * (time.perf_counter() - th)))
    torch.cuda.synchronize()
    gc.enable()
    step_ms = [a.elapsed_time(b) for a, b in evs]
    t_total_local = sum(step_ms)  # ms, device time of the K steps
    barrier()
    wall_ms = 1e3 * (time.perf_counter() - t_wall0)
    clocks = sampler.stop()
    ms_per_step = max_over_ranks(t_total_local) / K
    value = batch * world / (ms_per_step / 1e3)
    launches = K

    # size-independent checks at the full size (cheap, outside the timed region)
    n_ok = int((sol.result == 0).sum().item())
    steps_t = sol.stats["num_steps"]
    mean_steps = float(steps_t.float().mean().item())
    checks = {"successful_frac": n_ok / batch, "mean_num_steps": mean_steps, "max_num_steps": int(steps_t.max().item())}
    if "truth" in inp:
        err = (sol.value - torch.as_tensor(inp["truth"], device=dev)).abs().median(0).values
        checks["median_abs_err_vs_truth_max"] = float(err.max().item())

    # ---- e2e: host C entry point, pinned host buffers, copies inside the timed region (+ pageable, for an ordinary caller)
    e2e = None
    if do_e2e:
        inp_host = {k: torch.as_tensor(inp[k]).pin_memory().numpy() for k in keys}
        h2d = sum(inp_host[k].nbytes for k in keys)
        for _ in range(2):
            hs = run_host(name, inp_host, args.chunk)
        t_warm = time.perf_counter()
        while time.perf_counter() - t_warm < 0.03:
            hs = run_host(name, inp_host, args.chunk)
        barrier()
        t0 = time.perf_counter()
        for _ in range(K):
            hs = run_host(name, inp_host, args.chunk)
        torch.cuda.synchronize()
        e2e_ms = max_over_ranks(1e3 * (time.perf_counter() - t0)) / K
        d2h = hs.value.nbytes + hs.stats["num_steps"].nbytes + hs.result.nbytes
        checks["e2e_bitwise_equal_device_path"] = bool(np.array_equal(hs.value, sol.value.cpu().numpy()))
        e2e = {"value": batch * world / (e2e_ms / 1e3), "unit": "problems/s", "h2d_bytes_per_step": int(h2d) * world, "d2h_bytes_per_step": int(d2h) * world,
               "bytes_note": f"whole job: {world} rank(s) x ({int(h2d)} B in, {int(d2h)} B out) per step", "ms_per_step": e2e_ms, "api": "the *_host C entry point via the optimistix_b200 entry point called with NumPy inputs (page-locked)",
               "chunk": args.chunk}
        launches += K * ((batch + args.chunk - 1) // args.chunk)
        # the same call with ordinary (pageable) NumPy arrays: what a caller who did not pin anything gets
        inp_page = {k: np.array(inp[k], copy=True) for k in keys}
        run_host(name, inp_page, args.chunk)
        barrier()
        t0 = time.perf_counter()
        kp = max(1, min(K, 3))
        for _ in range(kp):
            run_host(name, inp_page, args.chunk)
        torch.cuda.synchronize()
        page_ms = max_over_ranks(1e3 * (time.perf_counter() - t0)) / kp
        e2e["pageable_inputs"] = {"value": batch * world / (page_ms / 1e3), "ms_per_step": page_ms}
        launches += (kp + 1) * ((batch + args.chunk - 1) // args.chunk)
        del inp_host, inp_page
        if name == "gauss2":
            # the same fits with the data rows stored as float16 (2^-11 relative quantisation, 20x below the noise of the data):
            # half the H2D bytes per problem; rows are converted exactly to fp32 on load (uint16 / int16 counts work the same way)
            d16 = inp["prob"].astype(np.float16)
            h16 = {"y0": torch.as_tensor(inp["y0"]).pin_memory().numpy(), "shared": inp["shared"], "prob": torch.as_tensor(d16).pin_memory().numpy()}
            for _ in range(2):
                hs16 = run_host(name, h16, args.chunk)
            barrier()
            t0 = time.perf_counter()
            for _ in range(K):
                hs16 = run_host(name, h16, args.chunk)
            torch.cuda.synchronize()
            ms16 = max_over_ranks(1e3 * (time.perf_counter() - t0)) / K
            dv16 = {"y0": torch.as_tensor(h16["y0"], device=dev), "shared": torch.as_tensor(inp["shared"], device=dev),
                    "prob": torch.as_tensor(d16, device=dev)}
            for _ in range(2):
                s16 = run_device(name, dv16)
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
            for _ in range(K):
                s16 = run_device(name, dv16)
            ev1.record()
            torch.cuda.synchronize()
            dms16 = max_over_ranks(ev0.elapsed_time(ev1)) / K
            e2e["float16_data"] = {"value": batch * world / (ms16 / 1e3), "ms_per_step": ms16, "h2d_bytes_per_step": int(h16["y0"].nbytes + h16["prob"].nbytes + inp["shared"].nbytes) * world,
                                  "device_resident_ms_per_step": dms16, "device_resident_value": batch * world / (dms16 / 1e3),
                                  "successful_frac": float((s16.result == 0).float().mean().item()), "mean_num_steps": float(s16.stats["num_steps"].float().mean().item()),
                                  "e2e_bitwise_equal_device_path": bool(np.array_equal(hs16.value, s16.value.cpu().numpy())),
                                  "note": "same curves, data rows stored as float16; NOT the headline configuration (fp32 data)"}
            launches += (K + 2) * ((batch + args.chunk - 1) // args.chunk) + K + 2
            del h16, dv16, d16

    cpu = None
    if do_cpu and world == 1 and rank == 0:
        sample = min(CPU_SAMPLE[name], batch)
        out, dt, cores, kind_note = cpu_reference(name, inp, sample)
        cpu = {"value": sample / dt, "unit": "problems/s", "cores": cores, "kind": "port",
               "sample": f"first {sample} problems of the same inputs, one pass; {kind_note}"}
        gv = sol.value[:sample].cpu().numpy()
        gs = sol.stats["num_steps"][:sample].cpu().numpy()
        tol = 1e-4 if dtype == "f32" else 1e-8
        scaled = np.max(np.abs(gv - out["value"]) / (tol + tol * np.abs(out["value"])), axis=1)
        checks["gpu_vs_cpu_port_within_tol_frac"] = float(np.mean(scaled <= 1.0))
        checks["gpu_vs_cpu_port_result_codes_equal_frac"] = float(np.mean(sol.result[:sample].cpu().numpy() == out["result"]))
        checks["gpu_vs_cpu_port_num_steps_equal_frac"] = float(np.mean(gs == out["num_steps"]))
        checks["cpu_port_mean_num_steps"] = float(out["num_steps"].mean())
        checks["gpu_mean_num_steps_same_sample"] = float(gs.mean())
        checks["mean_num_steps_bias_pct_vs_cpu_port"] = float(100.0 * (gs.mean() / out["num_steps"].mean() - 1.0))
    del inp_dev
    if rank != 0:
        return None

    # ---- roofline of the dominant (only) kernel
    peak, peak_src = _peaks()
    achieved = bytes_per_problem * batch / (ms_per_step / 1e3) / 1e9  # GB/s per GPU (per launch)
    tkey = {"gauss2": "lsq_solve_kernel<Gauss2<float>>", "newton": "root_solve_kernel<ImplicitEuler<double>>",
            "bfgs": "bfgs_solve_kernel<Rosenbrock<double>>"}.get(name, "")
    traffic, traffic_note = _traffic(tkey, batch)
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "traffic_note": traffic_note,
                "kernel": {"gauss2": "lsq_solve_kernel_p2", "exp_decay": "lsq_solve_kernel", "newton": "root_solve_kernel", "bfgs": "bfgs_solve_kernel"}[name],
                "algorithmic_bytes_per_problem": bytes_per_problem, "peak_source": peak_src,
                "note": "HBM is NOT the binding roof for this path (SURVEY F8): state is on-chip, ~0.5 Mflop per 1.1 KB; see 'fp'"}
    flops_per_iter = {"gauss2": 50e3, "exp_decay": 6.5e3, "newton": 28e3, "bfgs": 41e3}[name]
    fp_achieved = value / world * mean_steps * flops_per_iter / 1e12
    fp_peak, fp_src = _fp_peak(dtype)
    fp = {"achieved_tflops_algorithmic": fp_achieved, "peak_tflops": fp_peak, "peak_source": fp_src, "frac": fp_achieved / fp_peak,
          "note": "algorithmic flops/iteration from SURVEY 8d x measured mean iterations against the MEASURED FMA-pipe peak"}
    return {"metric": "LM problems solved/sec" if name in ("gauss2", "exp_decay") else f"{name} problems solved/sec",
            "value": value, "unit": "problems/s", "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_per_step,
            "higher_is_better": True, "dtype": dtype, "data": "synthetic", "batch_per_gpu": batch, "global_batch": batch * world,
            "workload": desc, "roofline": roofline, "fp": fp, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches,
            "clocks": clocks, "checks": checks, "wall_ms_timed_region": wall_ms, "ms_each_step": [round(v, 4) for v in step_ms],
            "host_us_each_step": host_us}


def bind_to_gpu_numa_node(local_rank):
    """Pin this process (and therefore the pages of the pinned staging buffers it allocates afterwards: first touch) to the
    CPUs of the NUMA node its GPU hangs off.  Without it the 8 ranks of one box pull their 1.1 GB per step out of whatever
    node the scheduler left them on (round 1: e2e efficiency 0.42 at 8 GPUs).  Best effort: silently keeps the old affinity."""
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        path = f"/sys/bus/pci/devices/{bus.lower()[-12:]}/local_cpulist"
        if not os.path.exists(path):
            path = f"/sys/bus/pci/devices/{bus.lower()}/local_cpulist"
        with open(path) as fh:
            spec = fh.read().strip()
        cpus = set()
        for part in spec.split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        if cpus:
            os.sched_setaffinity(0, cpus)
            return {"cpus": len(cpus), "cpulist": spec}
    except Exception as e:  # noqa: BLE001
        return {"error": repr(e)[:120]}
    return {"cpus": 0}


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
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="headline workload only")
    ap.add_argument("--no-numa", action="store_true", help="do not bind the process to the GPU's NUMA node")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    name = args.workload
    K = max(args.steps, 1)

    # ---------------------------------------------------------------- reference arm: CPU only, rank 0 only
    if args.impl == "reference":
        if rank != 0:
            return
        if name == "gram":
            print(json.dumps({"impl": "reference", "unavailable": "gram workload has no CPU reference arm (68.7 GB QR); see tests/test_gram_gpu.py for parity"}))
            return
        desc, batch, bytes_per_problem, dtype = WORKLOADS[name]
        if args.batch:
            batch = args.batch
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
                "config": {"workload": desc, "batch_per_gpu": batch, "global_batch": batch * args.gpus,
                           "problems_per_timed_step_of_this_arm": sample,
                           "note": "the CPU arm solves a bounded sample of the workload per step (the per-problem rate is what is compared); it does not grow with --gpus"},
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
    numa = None if args.no_numa else bind_to_gpu_numa_node(local_rank)
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import datetime

        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=600))
    W = max(args.warmup, 3)

    if name == "gram":
        line = bench_gram(args, rank, world, dev, K, W)
        if rank == 0:
            print(json.dumps(line))
        _shutdown(world)
        return

    desc, batch, bytes_per_problem, dtype = WORKLOADS[name]
    if args.batch:
        batch = args.batch
    inp = make_inputs(name, batch, rank)
    head = bench_batched(name, batch, inp, args, rank, world, dev, K, W, not args.no_e2e, not args.no_cpu)

    secondary = None
    if not args.no_secondary and name == "gauss2" and not args.batch:
        secondary = {}

        def guarded(key, fn):
            try:
                r = fn()
                if rank == 0:
                    secondary[key] = r
            except Exception as e:  # noqa: BLE001 - a failing secondary workload must not take the headline line with it
                if rank == 0:
                    secondary[key] = {"error": repr(e)[:300]}

        # configs[1] as BASELINE words it: 1 Mi fits IN TOTAL, batch-sharded over the ranks (strong scaling)
        if world > 1:
            bs = WORKLOADS["gauss2"][1] // world
            sl = {k: (v if k == "shared" else v[:bs]) for k, v in inp.items()}
            guarded("gauss2_strong", lambda: _tag(bench_batched("gauss2", bs, sl, args, rank, world, dev, K, W, not args.no_e2e, False), "strong"))
        else:
            secondary["gauss2_strong"] = "at 1 GPU the strong-scaled and the weak-scaled configuration are the same run: see the headline"
        del inp
        for nm in ("exp_decay", "newton", "bfgs"):
            tot = WORKLOADS[nm][1]
            bs = max(tot // world, 1)
            guarded(nm, lambda nm=nm, bs=bs: _tag(bench_batched(nm, bs, make_inputs(nm, bs, rank), args, rank, world, dev, K, W, not args.no_e2e,
                                                                not args.no_cpu), "strong"))
        guarded("gram", lambda: bench_gram(args, rank, world, dev, K, W))

    if rank == 0:
        line = {"metric": head["metric"], "value": head["value"], "unit": head["unit"], "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": dtype,
                "data": "synthetic",
                "config": {"workload": desc, "batch_per_gpu": batch, "global_batch": batch * world, "parallelism": f"batch-sharded x{world}, no collective",
                           "l2": "inputs (>= 1 GB at the headline size) exceed the 126 MB L2; every step re-reads them from HBM",
                           "numa_binding": numa},
                "roofline": head["roofline"], "fp": head["fp"], "cpu_baseline": head["cpu_baseline"], "e2e": head["e2e"],
                "gpu_launches": head["gpu_launches"], "clocks": head["clocks"], "checks": head["checks"],
                "wall_ms_timed_region": head["wall_ms_timed_region"], "ms_each_step": head["ms_each_step"],
                "host_us_each_step": head["host_us_each_step"]}
        if secondary is not None:
            line["secondary"] = secondary
        print(json.dumps(line))
    _shutdown(world)


def _tag(d, scaling):
    if d is not None:
        d["scaling"] = scaling
    return d


def _shutdown(world):
    if world > 1:
        import torch
        import torch.distributed as dist

        from optimistix_b200._gn_large import PeerExchange

        torch.cuda.synchronize()
        dist.barrier()
        PeerExchange.close_all()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
