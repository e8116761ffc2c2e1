"""GaussNewton on ONE large, row-sharded least-squares problem (BASELINE configs[4]; SURVEY sections 3.4 and 8e).

Each rank holds a contiguous shard of the rows of J (and of d); every rank holds the same y.  Per iteration:
    r      = J y_eval - d                       } b200opt_gn_gram_linear: ONE pass over J (TMA -> shared memory); the
    G,g,rr = J^T J, J^T r, r^T r                } tile feeds the tcgen05 tensor cores (fp32 accumulate in TMEM) and the
                                                  CUDA-core residual / J^T r / f-test statistic at the same time
                                                  (unfused pair: b200opt_linear_residual + b200opt_gn_gram)
    all-gather [G | g | rr | f-test statistic] over the ranks      -- the ONLY exchange: n^2+n+2 floats per rank
    p      = G^-1 g                             b200opt_gn_solve_parts: sums the gathered partials in rank order in fp64
                                                  (sum of G, g, rr; max of the statistic), then fp64 Cholesky; every rank
                                                  computes bit-identical values, so no second collective is needed
    y, y_eval = y_eval, y_eval - p
which is AbstractGaussNewton.step with search = LearningRate(1.0) (always accept) and descent = NewtonDescent
(optimistix/_solver/gauss_newton.py:243-334, :64-78, learning_rate.py:24-35) under the driver of
_iterate.py:194-268, with the normal equations in place of the reference's QR (north_star).

With `CudaOps` the whole loop runs on the DEVICE (`b200opt_gn_iterate`): the carry (y, y_eval, num_steps, terminate,
result) lives in a device state, an iteration is two kernels (Gram pass with the cross-CTA reduction and the publication
to the peers folded in; one-CTA solve that also applies the Cauchy test and advances the carry), frozen iterations are
no-ops, and the host enqueues `check_every` iterations ahead before it reads the state back.  The step-by-step host loop
below it is the same algorithm with the kernels injected through `ops`: it is what runs over NCCL when CUDA IPC is not
available, and what the CPU `gloo` test exercises with world_size > 1.
"""
from __future__ import annotations

import ctypes as C

from . import _lib
from ._solution import RESULTS, Solution

N = 256
OUT_FLOATS = N * N + N + 2  # [G | g | rr | f-stat bits]: one rank's contribution to the normal equations


class CudaOps:
    """The CUDA kernels behind the C ABI, on torch's current stream."""

    def __init__(self, device):
        import torch

        self.t = torch
        self.lib = _lib.load()
        self.device = device
        nbytes = self.lib.b200opt_gn_gram_workspace_size(N)
        self.ws = torch.empty(nbytes, dtype=torch.uint8, device=device)
        self.launches = 0

    def _stream(self):
        return self.t.cuda.current_stream(self.device).cuda_stream

    def residual(self, J, y, d, r_out, r_old, rtol, atol, fmax):
        fmax.zero_()
        _lib.check(self.lib.b200opt_linear_residual(J.data_ptr(), y.data_ptr(), d.data_ptr(), J.shape[0], J.shape[1], r_out.data_ptr(),
                                                    r_old.data_ptr() if r_old is not None else None, rtol, atol,
                                                    fmax.data_ptr(), self._stream()), "b200opt_linear_residual")
        self.launches += 1

    def gram(self, J, r, out):
        _lib.check(self.lib.b200opt_gn_gram(J.data_ptr(), r.data_ptr(), J.shape[0], J.shape[1], out.data_ptr(), self.ws.data_ptr(),
                                            self._stream()), "b200opt_gn_gram")
        self.launches += 2

    def residual_gram(self, J, y, d, r_out, r_old, rtol, atol, fmax, out):
        """Fused pass: r = J y - d, the f-space statistic, G, g and rr from ONE read of J (b200opt_gn_gram_linear)."""
        fmax.zero_()
        _lib.check(self.lib.b200opt_gn_gram_linear(J.data_ptr(), y.data_ptr(), d.data_ptr(),
                                                   r_old.data_ptr() if r_old is not None else None, r_out.data_ptr(),
                                                   fmax.data_ptr(), rtol, atol, J.shape[0], J.shape[1], out.data_ptr(), self.ws.data_ptr(),
                                                   self._stream()), "b200opt_gn_gram_linear")
        self.launches += 2

    def residual_gram_xchg(self, J, y, d, r_out, r_old, rtol, atol, fmax, out, xchg, parity, seq):
        """residual_gram + publication of this rank's block into every peer's exchange buffer (same two kernels)."""
        fmax.zero_()
        _lib.check(self.lib.b200opt_gn_gram_linear_xchg(J.data_ptr(), y.data_ptr(), d.data_ptr(),
                                                        r_old.data_ptr() if r_old is not None else None, r_out.data_ptr(),
                                                        fmax.data_ptr(), rtol, atol, J.shape[0], J.shape[1], out.data_ptr(), self.ws.data_ptr(),
                                                        xchg.bases.data_ptr(), xchg.world, xchg.rank, parity, seq, self._stream()),
                   "b200opt_gn_gram_linear_xchg")
        self.launches += 2

    def solve_xchg(self, xchg, parity, seq, lam, p, status, summary):
        _lib.check(self.lib.b200opt_gn_solve_xchg(xchg.local, xchg.world, parity, seq, N, float(lam), p.data_ptr(), status.data_ptr(),
                                                  summary.data_ptr(), self.ws.data_ptr(), self._stream()), "b200opt_gn_solve_xchg")
        self.launches += 1

    def solve(self, out, lam, p, status):
        _lib.check(self.lib.b200opt_gn_solve(out.data_ptr(), N, float(lam), p.data_ptr(), status.data_ptr(), self.ws.data_ptr(),
                                             self._stream()), "b200opt_gn_solve")
        self.launches += 1

    # ---- state-driven iteration (the whole loop on the device)
    def state_new(self):
        return self.t.empty(self.lib.b200opt_gn_state_size(), dtype=self.t.uint8, device=self.device)

    def state_init(self, state, y0, rtol, atol, max_steps, solver, timeout_ms, r_old0, out):
        _lib.check(self.lib.b200opt_gn_init(state.data_ptr(), y0.data_ptr(), int(y0.shape[0]), float(rtol), float(atol), int(max_steps), int(solver),
                                            int(timeout_ms), r_old0.data_ptr(), r_old0.shape[0], out.data_ptr(), self.ws.data_ptr(),
                                            self._stream()), "b200opt_gn_init")
        self.launches += 2

    def iterate(self, J, d, r0, r1, state, out, xchg, rtol, atol, lam=0.0, fold=False):
        bases = xchg.bases.data_ptr() if xchg is not None else None
        local = xchg.local if xchg is not None else None
        _lib.check(self.lib.b200opt_gn_iterate(J.data_ptr(), d.data_ptr(), r0.data_ptr(), r1.data_ptr(), J.shape[0], J.shape[1], state.data_ptr(),
                                               out.data_ptr(), self.ws.data_ptr(), bases, local, xchg.world if xchg is not None else 1,
                                               xchg.rank if xchg is not None else 0, float(rtol), float(atol), float(lam), int(bool(fold)),
                                               self._stream()),
                   "b200opt_gn_iterate")
        self.launches += 2 if fold else 3

    def solve_parts(self, parts, nparts, stride, lam, p, status, summary):
        """parts: the all-gathered [G | g | rr | f-stat] blocks of every rank (or this rank's own block, nparts = 1)."""
        _lib.check(self.lib.b200opt_gn_solve_parts(parts.data_ptr(), int(nparts), int(stride), N, float(lam), p.data_ptr(),
                                                   status.data_ptr(), summary.data_ptr(), self.ws.data_ptr(), self._stream()),
                   "b200opt_gn_solve_parts")
        self.launches += 1


class PeerExchange:
    """The collective of the row-sharded problem done by our own kernels over peer memory (NVLink / NVSwitch): every rank
    owns one exchange buffer, mapped into all peers with CUDA IPC; `b200opt_gn_gram_linear_xchg` stores this rank's
    [G | g | rr | f-stat] block into every peer's buffer and raises a flag, `b200opt_gn_solve_xchg` waits for all flags
    and sums the blocks in rank order.  torch.distributed is used ONCE, to swap the 64-byte IPC handles."""

    def __init__(self, device, group=None):
        import torch
        import torch.distributed as dist

        self.lib = _lib.load()
        self.t = torch
        self.device = device
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        nbytes = self.lib.b200opt_xchg_bytes(self.world, N)
        if nbytes <= 0:
            raise _lib.B200OptError("b200opt_xchg_bytes failed")
        ptr = C.c_void_p()
        handle = C.create_string_buffer(64)
        with torch.cuda.device(device):
            _lib.check(self.lib.b200opt_xchg_alloc(nbytes, C.byref(ptr), handle), "b200opt_xchg_alloc")
        self.local = ptr.value
        handles = [None] * self.world
        dist.all_gather_object(handles, handle.raw, group=group)
        self.peers = []
        bases = []
        with torch.cuda.device(device):
            for q in range(self.world):
                if q == self.rank:
                    bases.append(self.local)
                    continue
                pp = C.c_void_p()
                _lib.check(self.lib.b200opt_xchg_open(handles[q], C.byref(pp)), "b200opt_xchg_open")
                self.peers.append(pp.value)
                bases.append(pp.value)
        self.bases = torch.tensor(bases, dtype=torch.int64, device=device)
        self.seq = 0
        dist.barrier(group=group)  # every buffer is zeroed and mapped before anybody publishes

    def next(self):
        self.seq += 1
        return self.seq & 1, self.seq

    _cache = {}

    @classmethod
    def shared(cls, device, group=None):
        """The exchange of (device, group), created on first use and REUSED by every later solve (IPC mappings, the
        all_gather_object of the handles and the barrier are paid once, nothing leaks per call).  Collective."""
        import torch

        key = (torch.device(device).index, group)  # the group object itself: the cache keeps it alive, an id() could be recycled
        x = cls._cache.get(key)
        if x is None or x.local is None:
            x = cls(device, group)
            cls._cache[key] = x
        return x

    @classmethod
    def close_all(cls):
        for x in cls._cache.values():
            x.close()
        cls._cache.clear()

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001 - interpreter teardown: the CUDA context may be gone
            pass

    def close(self):
        for pp in self.peers:
            self.lib.b200opt_xchg_close(pp)
        self.peers = []
        if self.local:
            self.lib.b200opt_xchg_free(self.local)
            self.local = None


STATE_TAIL_INTS = 12  # int32 words after y[256], y_eval[256] in the device state (include/b200opt.h)


def _device_loop(ops, J, d, y, solver, max_steps, throw, xchg, check_every, linear_solver, timeout_ms, fold_reduce=False):
    """The loop of _iterate.py:194-268 on the device: enqueue `check_every` iterations, read the 48-byte tail of the state."""
    import torch

    dev = J.device
    m_local = J.shape[0]
    rtol, atol = float(solver.rtol), float(solver.atol)
    r0 = torch.empty(m_local, dtype=torch.float32, device=dev)
    r1 = torch.empty(m_local, dtype=torch.float32, device=dev)
    out = torch.empty(OUT_FLOATS, dtype=torch.float32, device=dev)
    state = ops.state_new()
    ops.state_init(state, y, rtol, atol, min(max_steps, 2**31 - 1), linear_solver, timeout_ms, r1, out)
    tail = state[2 * N * 4:].view(torch.int32)
    tail_host = torch.empty(STATE_TAIL_INTS, dtype=torch.int32, pin_memory=True)
    done_ev = torch.cuda.Event()
    steps = terminate = 0
    result = RESULTS.successful
    enq = 0
    while True:
        tail_host.copy_(tail, non_blocking=True)
        done_ev.record()
        done_ev.synchronize()
        steps, terminate, result = (int(v) for v in tail_host[:3])
        if terminate or result != RESULTS.successful or steps >= max_steps:
            break
        k = min(check_every, max_steps - steps)
        for _ in range(k):
            ops.iterate(J, d, r0, r1, state, out, xchg, rtol, atol, fold=fold_reduce)
        enq += k
    if result < 0:
        raise _lib.B200OptError("peer exchange timed out: a rank never published its block of the normal equations")
    if result == RESULTS.successful and not terminate:
        result = RESULTS.nonlinear_max_steps_reached  # _iterate.py:253-257
    if throw and result != RESULTS.successful:
        raise RuntimeError(f"{RESULTS.name(result)}: {RESULTS[result]}")
    fl = tail_host.view(torch.float32)
    value = state[: N * 4].view(torch.float32)[: J.shape[1]].clone()  # the state is laid out 256 wide; unknowns beyond n are padding
    stats = {"num_steps": steps, "max_steps": max_steps, "loss": 0.5 * float(fl[6]), "gpu_launches": getattr(ops, "launches", None),
             "iterations_enqueued": enq, "fp64_fallbacks": int(tail_host[10]),
             "exchange": "peer memory (device loop)" if xchg is not None else "single rank (device loop)"}
    return Solution(value=value, result=result, aux=None, stats=stats, state={"gram": out, "residual": r0 if steps % 2 else r1})


def gauss_newton_large(J, d, y0, solver, *, max_steps=256, throw=True, group=None, ops=None, fused=True, exchange="peer",
                       device_loop=True, check_every=4, linear_solver=0, timeout_ms=0, fold_reduce=False):
    """Solve min_y |J y - d|^2 with this rank's row shard (J [m_local, 256] fp32, d [m_local]); returns a Solution.

    `group`: a torch.distributed process group (None = the default group if initialised, else single rank).
    `fused`: form the residual inside the Gram kernel (J is read once per iteration) when `ops` offers it.
    `exchange`: "peer" = the per-iteration exchange is done by our kernels over peer memory (PeerExchange; falls back to
    "nccl" = one all_gather_into_tensor per iteration when CUDA IPC is not available or `ops` has no kernels for it).
    `device_loop`: run the loop on the device (b200opt_gn_iterate; single rank or peer exchange), the host reading the
    state every `check_every` iterations; False = the step-by-step host loop (one host read per iteration).
    `linear_solver`: 0 = equilibrated fp32 blocked Cholesky + fp64 refinement (fp64 fallback), 1 = fp64 Cholesky.
    `fold_reduce`: reduce and publish inside the Gram kernel (one cooperative launch) instead of a separate small kernel;
    same bits either way."""
    import torch
    import torch.distributed as dist

    if getattr(solver.norm, "norm_name", "max") != "max":
        raise NotImplementedError("the large Gauss-Newton path implements max_norm termination")
    if _lib.linear_solver_id(getattr(getattr(solver, "descent", None), "linear_solver", None)) != _lib.LINSOLVE_NORMAL_CHOLESKY:
        raise NotImplementedError("linear_rows (m x 256 with m ~ 6.7e7): the row-sharded path forms the normal equations on the tensor "
                                  "cores; a tall-skinny QR across ranks is not implemented - use linear_solver=NormalCholesky()")
    distributed = dist.is_available() and dist.is_initialized() and (group is not None or dist.get_world_size() > 1)
    dev = J.device
    if ops is None:
        if not J.is_cuda:
            raise _lib.B200OptError("the large Gauss-Newton path needs CUDA tensors (there is no CPU fallback)")
        ops = CudaOps(dev)
    n = int(J.shape[1]) if J.ndim == 2 else -1
    if J.dtype != torch.float32 or J.ndim != 2 or not (4 <= n <= N) or n % 4 or not J.is_contiguous():
        raise ValueError(f"J must be a contiguous float32 [m_local, n] tensor with 4 <= n <= {N} and n % 4 == 0 (the kernels work on the "
                         f"matrix padded with zero columns to {N}; the padding is never read from memory)")
    m_local = J.shape[0]
    d = d.to(torch.float32).contiguous()
    y = y0.to(device=dev, dtype=torch.float32).contiguous().clone()
    if y.shape != (n,):
        raise ValueError(f"y0 must have shape [{n}] (the number of columns of J)")
    rtol, atol = float(solver.rtol), float(solver.atol)
    fused = fused and hasattr(ops, "residual_gram")
    y_unpadded = y
    if n < N:  # the step-by-step kernels read y 256 wide; the padding stays exactly zero (unit diagonal, zero gradient)
        y = torch.zeros(N, dtype=torch.float32, device=dev)
        y[:n] = y_unpadded
    y_eval = y.clone()
    r_old = torch.zeros(m_local, dtype=torch.float32, device=dev)  # gauss_newton.py:229-230: zero-filled f_info
    r_new = torch.empty_like(r_old)
    out = torch.empty(OUT_FLOATS, dtype=torch.float32, device=dev)  # [G | g | rr | f-stat]: what one rank contributes
    fmax = out[N * N + N + 1:].view(torch.int32)  # IEEE bits of a non-negative float: integer max == float max
    world = dist.get_world_size(group) if distributed else 1
    xchg = None
    if distributed and exchange == "peer" and fused and hasattr(ops, "residual_gram_xchg"):
        # collective decision: either every rank maps every peer, or all fall back to NCCL.  The exchange is cached per
        # (device, group): later solves reuse the mappings (and skip this handshake's allocations)
        try:
            # (the step-by-step host loop numbers its exchanges on the host: it gets a private buffer, closed on return)
            xchg = PeerExchange.shared(dev, group) if device_loop else PeerExchange(dev, group)
            ok = torch.ones(1, dtype=torch.int32, device=dev)
        except Exception as e:  # noqa: BLE001
            import warnings

            warnings.warn(f"peer-memory exchange unavailable ({e}); using the NCCL all-gather")
            xchg, ok = None, torch.zeros(1, dtype=torch.int32, device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
        if int(ok.item()) == 0:
            if xchg is not None and not device_loop:
                xchg.close()
            xchg = None
    if device_loop and fused and hasattr(ops, "iterate") and (not distributed or xchg is not None):
        with torch.cuda.device(dev):
            return _device_loop(ops, J, d, y_unpadded, solver, max_steps, throw, xchg, max(1, int(check_every)), int(linear_solver), int(timeout_ms),
                                bool(fold_reduce))
    try:
        gathered = torch.empty(world * OUT_FLOATS, dtype=torch.float32, device=dev) if (distributed and xchg is None) else out
        summary = torch.zeros(2, dtype=torch.float32, device=dev)  # total r^T r, max f-stat
        p = torch.empty(N, dtype=torch.float32, device=dev)
        status = torch.zeros(1, dtype=torch.int32, device=dev)
        result = RESULTS.successful if bool(torch.isfinite(y).all()) else RESULTS.nonfinite  # _iterate.py:216-219
        terminate = False
        steps = 0
        while not terminate and result == RESULTS.successful and steps < max_steps:
            y_fin_t = torch.isfinite(y).all()  # _iterate.py:216,238: finiteness of the carry's y from BEFORE this step
            if xchg is not None:
                parity, seq = xchg.next()
                ops.residual_gram_xchg(J, y_eval, d, r_new, r_old, rtol, atol, fmax, out, xchg, parity, seq)
            elif fused:
                ops.residual_gram(J, y_eval, d, r_new, r_old, rtol, atol, fmax, out)  # one read of J per iteration
            else:
                ops.residual(J, y_eval, d, r_new, r_old, rtol, atol, fmax)  # gauss_newton.py:253 + f-space half of :280-288
                ops.gram(J, r_new, out)
            if xchg is not None:
                ops.solve_xchg(xchg, parity, seq, 0.0, p, status, summary)  # waits for every rank's flag, then as below
            else:
                if distributed:
                    dist.all_gather_into_tensor(gathered, out, group=group)  # the only exchange of the iteration
                ops.solve_parts(gathered, world, OUT_FLOATS, 0.0, p, status, summary)  # NewtonDescent.query: newton = G^-1 g
            y_diff = y_eval - y
            y_conv_t = (y_diff.abs() / (atol + rtol * y_eval.abs())).max() < 1  # _misc.py:306-308
            f_conv_t = summary[1] < 1  # NaN compares false, like jnp.max
            # ONE host read per iteration (the flags of this step and the linear solve's result code)
            y_old_finite, y_conv, f_conv, result = torch.stack(
                [y_fin_t.to(torch.int32), y_conv_t.to(torch.int32), f_conv_t.to(torch.int32), status[0]]).tolist()
            terminate = bool(y_conv) and bool(f_conv)  # LearningRate always accepts (learning_rate.py:24-35)
            y = y_eval
            y_eval = y - p  # gauss_newton.py:126-127 with step_size 1
            r_old, r_new = r_new, r_old
            # `result` is the linear solve's (gauss_newton.py:316-318)
            if result < 0:
                raise _lib.B200OptError("peer exchange timed out: a rank never published its block of the normal equations")
            if not y_old_finite:
                result = RESULTS.nonfinite
            steps += 1
    finally:
        if xchg is not None:  # private exchange of the host loop: nobody may still be storing into a buffer that is freed
            torch.cuda.synchronize(dev)
            dist.barrier(group=group)
            xchg.close()
    if result == RESULTS.successful and not terminate:
        result = RESULTS.nonlinear_max_steps_reached  # _iterate.py:253-257
    if throw and result != RESULTS.successful:
        raise RuntimeError(f"{RESULTS.name(result)}: {RESULTS[result]}")
    stats = {"num_steps": steps, "max_steps": max_steps, "loss": 0.5 * float(summary[0].item()),
             "gpu_launches": getattr(ops, "launches", None),
             "exchange": ("peer memory (host loop)" if xchg is not None else "nccl all-gather (host loop)") if distributed else "single rank (host loop)"}
    return Solution(value=y[:n], result=result, aux=None, stats=stats, state={"gram": out, "residual": r_old})
