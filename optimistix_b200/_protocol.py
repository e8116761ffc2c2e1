"""The solver protocol of the reference - `init / step / terminate / postprocess` (optimistix/_iterate.py:33-184) - on the
CUDA step kernels.

The reference's entry points are a `while` loop around these four methods (`_iterate.py:194-268`); a user may also call
them directly to step through a solve (docs/examples/interactive.ipynb).  Here one `step` is: torch evaluates the user's
function and its derivatives at the trial point (the role of `jax.linearize`), then ONE launch of the per-iteration step
kernel (`b200opt_lsq_step`, `b200opt_root_step`, `b200opt_bfgs_step`) does everything else of the reference's `step` for
every problem of the batch: search, accept / reject, descent, termination test, result codes.  The entry points
(`least_squares`, `root_find`, `minimise` with a Python callable) run exactly these sessions in a loop (`_generic.py`), so
stepping by hand and calling the entry point give bit-identical iterates.

Differences from the reference, by construction:
  * the batch is explicit (`y` is `[n]` or `[batch, n]`); `done` / `result` are device tensors with one entry per problem;
    a finished problem is frozen inside the kernel, like a finished lane of `jax.vmap(while_loop)`;
  * the solver state is a mutable object that owns device buffers: `step` updates it in place and returns it (the
    reference returns a new immutable pytree).  Its public attributes follow the reference's state classes
    (`_GaussNewtonState` gauss_newton.py:190-200, `_NewtonChordState` newton_chord.py:40-47, `_QuasiNewtonState`
    quasi_newton.py:155-169): `first_step`, `y_eval`, `f_info` (a `FunctionInfo`), `terminate`, `result`,
    `num_accepted_steps`, `search_state` (the step size);
  * `f_struct` / `aux_struct` are accepted and ignored (shapes are read off the first evaluation).
"""
from __future__ import annotations

import ctypes as C

from . import _lib
from ._function_info import FunctionInfo
from ._solution import RESULTS

_UNBOUNDED = 2**31 - 1


def _torch():
    import torch

    return torch


def _prep_y0(y0):
    t = _torch()
    y0 = t.as_tensor(y0)
    if not y0.is_cuda:
        if not t.cuda.is_available():
            raise _lib.B200OptError("the generic path needs a CUDA device (there is no CPU fallback)")
        y0 = y0.cuda()
    if not y0.dtype.is_floating_point:
        y0 = y0.to(t.get_default_dtype())
    if y0.dtype not in (t.float32, t.float64):
        raise TypeError(f"unsupported dtype {y0.dtype}")
    single = y0.ndim == 1
    if y0.ndim not in (1, 2):
        raise ValueError("y0 must have shape [n] (one problem) or [batch, n]")
    Y0 = (y0[None] if single else y0).contiguous()
    return Y0, single, (_lib.F32 if y0.dtype == t.float32 else _lib.F64)


def _detached(args):
    """args with every tensor leaf detached: the forward iteration is never recorded by autograd (ImplicitAdjoint
    differentiates the optimality condition at the solution instead, _adjoint.py)."""
    t = _torch()
    from torch.utils._pytree import tree_map

    return tree_map(lambda x: x.detach() if isinstance(x, t.Tensor) else x, args)


def _vmapped(f, args, args_in_dims):
    """vmap f(y, args) over the batch axis of y (and of the leaves of args selected by args_in_dims)."""
    t = _torch()
    if args_in_dims is None:
        return t.func.vmap(lambda y: f(y, args))
    vm = t.func.vmap(f, in_dims=(0, args_in_dims))
    return lambda Y: vm(Y, args)


def _value_and_jac(one, jac_mode):
    """y, a -> (value, jacobian) with ONE evaluation of the user's function per pass (the value rides along as `aux`)."""
    t = _torch()
    jacf = t.func.jacfwd if jac_mode == "fwd" else t.func.jacrev

    def both(y, a):
        def value_and_aux(z):
            v = one(z, a)
            return v, v

        J, v = jacf(value_and_aux, has_aux=True)(y)
        return v, J

    return both


class _Session:
    """Buffers and the per-iteration body shared by the three families."""

    kind = ""

    def _common(self, solver, y0, args, args_in_dims, max_steps):
        t = _torch()
        self.t = t
        self.lib = _lib.load()
        self.solver = solver
        self.Y0, self.single, self.dtype_code = _prep_y0(y0)
        self.B, self.n = self.Y0.shape
        self.dev = self.Y0.device
        self.args = args
        self.args_fwd = _detached(args)
        self.args_in_dims = args_in_dims
        self.max_steps_k = _UNBOUNDED if max_steps is None else int(max_steps)
        self.es = 4 if self.dtype_code == _lib.F32 else 8
        self.steps = t.empty((self.B,), dtype=t.int32, device=self.dev)
        self.res = t.empty((self.B,), dtype=t.int32, device=self.dev)
        self.active = t.zeros((1,), dtype=t.int32, device=self.dev)
        self.launches = 0
        self._hold = {}
        self.verbose = getattr(solver, "verbose", None)  # None, or a callable receiving name=(label, value) once per step

    def _stream(self):
        return self.t.cuda.current_stream(self.dev).cuda_stream  # the capture stream while a graph is being recorded

    # ---- what the protocol methods hand out
    def _sq(self, x):
        return x[0] if self.single else x

    @property
    def result(self):
        return self._sq(self.res)

    @property
    def num_steps(self):
        return self._sq(self.steps)

    @property
    def done(self):
        """terminate | failed: one flag per problem (`terminate` of the reference returns `(state.terminate, state.result)` and
        the driver also stops on a non-successful result, _iterate.py:230-232)."""
        return self._sq(self.terminate_flags() | (self.res != RESULTS.successful))

    @property
    def terminate(self):
        return self._sq(self.terminate_flags())

    def step(self):
        self.body()
        self.launches += 1
        return self


# ------------------------------------------------------------------------------------------ LM / GN
class LsqSession(_Session):
    """AbstractGaussNewton (gauss_newton.py:203-334) for an arbitrary residual function: state of `b200opt_lsq_step`."""

    kind = "lstsq"

    def __init__(self, fn, solver, y0, args, options=None, args_in_dims=None, max_steps=None, root_near_zero=False):
        from ._solvers import LevenbergMarquardt

        self._common(solver, y0, args, args_in_dims, max_steps)
        t, lib = self.t, self.lib
        if self.n > 64:
            raise NotImplementedError("generic least-squares step kernels support n <= 64 parameters (n <= 16: Gram in registers, one warp "
                                      "per problem; n <= 64: Gram in shared memory, one CTA per problem)")
        self.empty = self.B == 0
        if self.empty:
            return
        jac_mode = (options or {}).get("jac", "fwd")
        rj = _vmapped(_value_and_jac(lambda y, a: fn(y, a).reshape(-1), jac_mode), self.args_fwd, args_in_dims)
        self.rj = rj
        r, J = rj(self.Y0)
        self.m = int(r.shape[1])
        self._first = (r, J)
        stride = lib.b200opt_lsq_step_state_size(self.dtype_code, self.n, self.m)
        if stride <= 0:
            _lib.check(-1, "b200opt_lsq_step_state_size")
        self.state = t.empty((self.B, stride // self.es), dtype=self.Y0.dtype, device=self.dev)
        self.y = t.empty_like(self.Y0)
        self.y_eval_buf = t.empty_like(self.Y0)
        self.solver_id = _lib.SOLVER_LM if isinstance(solver, LevenbergMarquardt) else _lib.SOLVER_GN
        self.opts = _lib.make_options(solver.rtol, solver.atol, solver.norm, self.max_steps_k, root_near_zero=root_near_zero,
                                      linear_solver=getattr(getattr(solver, "descent", None), "linear_solver", None))
        _lib.check(lib.b200opt_lsq_step_init(self.dtype_code, self.B, self.n, self.m, self.Y0.data_ptr(), self.y.data_ptr(),
                                             self.y_eval_buf.data_ptr(), self.state.data_ptr(), self.steps.data_ptr(), self.res.data_ptr(),
                                             self._stream()), "b200opt_lsq_step_init")
        # layout of the carry (csrc/lsq_step.cu): [G packed lower | g | rr | delta | first | terminate | num_accepted | r_accepted]
        # csrc/lsq_step.cu pad_n: register kernel for n padded to 4 / 8 / 16 (16: fp32 only), 0 = the CTA kernel
        self.np_ = 4 if self.n <= 4 else 8 if self.n <= 8 else 16 if (self.n <= 16 and self.dtype_code == _lib.F32) else 0

    def body(self):
        rr, JJ = self._first if self._first is not None else self.rj(self.y_eval_buf)
        self._first = None
        rr, JJ = rr.contiguous(), JJ.contiguous()
        if self.verbose is not None:  # gauss_newton.py:298-312: the carry as it is when the step starts
            pre = dict(num_steps=("Step", self.num_steps.clone()), accepted_steps=("Accepted steps", self.num_accepted_steps.clone()),
                       loss_this_step=("Loss on this step", self._sq(0.5 * (rr * rr).sum(1))),
                       loss_last_accepted_step=("Loss on the last accepted step", 0.5 * self.f_info.sum_squares.clone()),
                       y=("y", self.y_eval.clone()))
        _lib.check(self.lib.b200opt_lsq_step(self.solver_id, self.dtype_code, self.B, self.n, self.m, rr.data_ptr(), JJ.data_ptr(),
                                             self.y.data_ptr(), self.y_eval_buf.data_ptr(), self.state.data_ptr(), C.byref(self.opts),
                                             self.steps.data_ptr(), self.res.data_ptr(), self.active.data_ptr(), self._stream()),
                   "b200opt_lsq_step")
        self._hold["in"] = (rr, JJ)  # the eager path must keep its inputs alive until the kernel has run
        if self.verbose is not None:
            self.verbose(**pre, step_size=("Step size", self.search_state), y_last_accepted_step=("y on the last accepted step", self._sq(self.y)))

    def finish(self):
        _lib.check(self.lib.b200opt_lsq_step_finish(self.dtype_code, self.B, self.n, self.m, self.state.data_ptr(), self.res.data_ptr(),
                                                    self._stream()), "b200opt_lsq_step_finish")

    # ---- the reference's _GaussNewtonState fields, read from the device carry
    def _head(self):
        if self.np_:
            ng = self.np_ * (self.np_ + 1) // 2
            return ng, self.np_
        return self.n * self.n, self.n  # CTA kernel (16 < n <= 64): full G, then g (csrc/lsq_step.cu big_state_head)

    def terminate_flags(self):
        ng, npad = self._head()
        return self.state[:, ng + npad + 3] != 0

    @property
    def first_step(self):
        ng, npad = self._head()
        return self._sq(self.state[:, ng + npad + 2] != 0)

    @property
    def search_state(self):
        """step size of the search: the trust-region radius (LM) / the learning rate (GN)."""
        ng, npad = self._head()
        return self._sq(self.state[:, ng + npad + 1])

    @property
    def num_accepted_steps(self):
        ng, npad = self._head()
        return self._sq(self.state[:, ng + npad + 4].to(self.t.int32))

    @property
    def y_eval(self):
        return self._sq(self.y_eval_buf)

    @property
    def f_info(self):
        """FunctionInfo at the accepted point, in the form the kernels carry (ResidualGram: r^T r, J^T J, J^T r)."""
        t = self.t
        ng, npad = self._head()
        n = self.n
        if self.np_:
            idx = t.tril_indices(npad, npad, device=self.dev)
            G = t.zeros((self.B, npad, npad), dtype=self.state.dtype, device=self.dev)
            G[:, idx[0], idx[1]] = self.state[:, :ng]
            G = G + G.transpose(1, 2) - t.diag_embed(t.diagonal(G, dim1=1, dim2=2))
            G = G[:, :n, :n]
        else:
            G = self.state[:, :ng].reshape(self.B, n, n)
        info = FunctionInfo.ResidualGram(sum_squares=self._sq(self.state[:, ng + npad]), gram=self._sq(G),
                                         grad=self._sq(self.state[:, ng : ng + n]))
        return info


# ------------------------------------------------------------------------------------------ Newton / Chord
class RootSession(_Session):
    """_AbstractNewtonChord (newton_chord.py:59-196) for an arbitrary `fn(y, args) -> [n]`: state of `b200opt_root_step`."""

    kind = "root"

    def __init__(self, fn, solver, y0, args, options=None, args_in_dims=None, max_steps=None):
        from ._solvers import Chord

        self._common(solver, y0, args, args_in_dims, max_steps)
        t, lib = self.t, self.lib
        n, dev = self.n, self.dev
        nmax = 160 if self.dtype_code == _lib.F64 else 224
        if n > nmax:
            raise NotImplementedError(f"generic Newton/Chord step kernels support n <= {nmax} unknowns in this dtype (n <= 32: one warp per "
                                      "problem, row-per-lane LU; above: one CTA per problem, the Jacobian in shared memory)")
        self.empty = self.B == 0
        if self.empty:
            return

        def one(y, a):
            out = fn(y, a).reshape(-1)
            if out.shape[0] != n:
                raise ValueError("root_find: fn must map [n] -> [n]")
            return out

        self.f_only = _vmapped(one, self.args_fwd, args_in_dims)
        self.f_and_J = _vmapped(_value_and_jac(one, "fwd"), self.args_fwd, args_in_dims)
        self.lower = self.upper = None
        if options:
            if options.get("lower") is not None:
                self.lower = t.broadcast_to(t.as_tensor(options["lower"], dtype=self.Y0.dtype, device=dev), (n,)).contiguous()
            if options.get("upper") is not None:
                self.upper = t.broadcast_to(t.as_tensor(options["upper"], dtype=self.Y0.dtype, device=dev), (n,)).contiguous()
        stride = lib.b200opt_root_step_state_size(self.dtype_code, n)
        if stride <= 0:
            _lib.check(-1, "b200opt_root_step_state_size")
        self.state = t.empty((self.B, stride // self.es), dtype=self.Y0.dtype, device=dev)
        self.y = t.empty_like(self.Y0)
        self.opts = _lib.make_options(solver.rtol, solver.atol, solver.norm, self.max_steps_k, cauchy_termination=solver.cauchy_termination,
                                      kappa=solver.kappa)
        _lib.check(lib.b200opt_root_step_init(self.dtype_code, self.B, n, self.Y0.data_ptr(), self.y.data_ptr(), self.state.data_ptr(),
                                              self.steps.data_ptr(), self.res.data_ptr(), self._stream()), "b200opt_root_step_init")
        self.chord = isinstance(solver, Chord)
        self._started = False

    def _launch(self, f, J):
        f, J = f.contiguous(), J.contiguous()
        _lib.check(self.lib.b200opt_root_step(self.dtype_code, self.B, self.n, f.data_ptr(), J.data_ptr(),
                                              self.lower.data_ptr() if self.lower is not None else None,
                                              self.upper.data_ptr() if self.upper is not None else None, self.y.data_ptr(),
                                              self.state.data_ptr(), C.byref(self.opts), self.steps.data_ptr(), self.res.data_ptr(),
                                              self.active.data_ptr(), self._stream()), "b200opt_root_step")
        self._hold["fJ"] = (f, J)

    def first_body(self):
        f, J = self.f_and_J(self.y)
        if self.chord:
            self._hold["J_fixed"] = J.contiguous()  # newton_chord.py:80-87: linearised once, at y0
        self._launch(f, J)
        self._started = True

    def body(self):
        if not self._started:
            return self.first_body()
        if self.chord:
            self._launch(self.f_only(self.y), self._hold["J_fixed"])
        else:
            self._launch(*self.f_and_J(self.y))

    def finish(self):
        _lib.check(self.lib.b200opt_root_step_finish(self.dtype_code, self.B, self.n, self.state.data_ptr(), self.res.data_ptr(),
                                                     self._stream()), "b200opt_root_step_finish")

    def terminate_flags(self):
        return self.state[:, 2] != 0  # csrc/generic_step.cu: [diffsize | diffsize_prev | terminate | step]

    @property
    def f_info(self):
        """FunctionInfo.Eval of the last evaluated point (the reference's Newton state keeps `f`)."""
        fJ = self._hold.get("fJ")
        return FunctionInfo.Eval(self._sq(fJ[0])) if fJ is not None else None


# ------------------------------------------------------------------------------------------ BFGS / DFP
class MinSession(_Session):
    """AbstractQuasiNewton (quasi_newton.py:176-300) for an arbitrary scalar function: state of `b200opt_bfgs_step`."""

    kind = "min"

    def __init__(self, fn, solver, y0, args, options=None, args_in_dims=None, max_steps=None, root_fn=None, autodiff_mode=None):
        self._common(solver, y0, args, args_in_dims, max_steps)
        t, lib = self.t, self.lib
        n, dev = self.n, self.dev
        self.empty = self.B == 0
        if self.empty:
            return
        if autodiff_mode is None:
            autodiff_mode = (options or {}).get("autodiff_mode", "bwd")

        def one(y, a):
            out = fn(y, a)
            if out.numel() != 1:
                raise ValueError("minimise function must output a single scalar.")  # _minimise.py:97-104
            return out.reshape(())

        if autodiff_mode == "fwd":  # options["autodiff_mode"] (_misc.py:245-258): gradient by forward-mode JVPs
            vj = _value_and_jac(one, "fwd")
            self.gv = _vmapped(lambda y, a: tuple(reversed(vj(y, a))), self.args_fwd, args_in_dims)
        else:
            self.gv = _vmapped(lambda y, a: t.func.grad_and_value(lambda z: one(z, a))(y), self.args_fwd, args_in_dims)
        self.rootv = None
        if root_fn is not None:
            norm = solver.norm
            self.rootv = _vmapped(lambda y, a: norm(root_fn(y, a).reshape(-1)), self.args_fwd, args_in_dims)
        nbytes = lib.b200opt_bfgs_step_state_size(self.dtype_code, n)
        if nbytes <= 0:
            _lib.check(-1, "b200opt_bfgs_step_state_size")
        self.state = t.empty((self.B, nbytes // self.es), dtype=self.Y0.dtype, device=dev)
        self.y = t.empty_like(self.Y0)
        self.y_eval_buf = t.empty_like(self.Y0)
        self.opts = _lib.make_options(solver.rtol, solver.atol, solver.norm, self.max_steps_k)
        self.update = _lib.UPDATE_DFP if getattr(solver, "_update", "bfgs") == "dfp" else _lib.UPDATE_BFGS
        self.use_inverse = 1 if getattr(solver, "use_inverse", True) else 0
        if not self.use_inverse and n > 64:
            raise NotImplementedError("use_inverse=False factorises the Hessian approximation in shared memory: n <= 64")
        _lib.check(lib.b200opt_bfgs_step_init(self.dtype_code, self.B, n, self.Y0.data_ptr(), self.y.data_ptr(), self.y_eval_buf.data_ptr(),
                                              self.state.data_ptr(), self.steps.data_ptr(), self.res.data_ptr(), self._stream()),
                   "b200opt_bfgs_step_init")

    def body(self):
        g, f = self.gv(self.y_eval_buf)
        g, f = g.contiguous(), f.contiguous()
        rn = self.rootv(self.y).to(self.Y0.dtype).contiguous() if self.rootv is not None else None
        if self.verbose is not None:  # quasi_newton.py:272-278
            n = self.n
            pre = dict(loss_this_step=("Loss on this step", self._sq(f.clone())),
                       loss_last_accepted_step=("Loss on the last accepted step", self._sq(self.state[:, n * n + 2 * n].clone())),
                       y=("y", self.y_eval.clone()))
        _lib.check(self.lib.b200opt_bfgs_step(self.update, self.use_inverse, self.dtype_code, self.B, self.n, f.data_ptr(), g.data_ptr(),
                                              rn.data_ptr() if rn is not None else None, self.y.data_ptr(), self.y_eval_buf.data_ptr(),
                                              self.state.data_ptr(), C.byref(self.opts), self.steps.data_ptr(), self.res.data_ptr(),
                                              self.active.data_ptr(), self._stream()), "b200opt_bfgs_step")
        self._hold["in"] = (g, f, rn)
        if self.verbose is not None:
            self.verbose(**pre, step_size=("Step size", self.search_state), y_last_accepted_step=("y on the last accepted step", self._sq(self.y)))

    def finish(self):
        _lib.check(self.lib.b200opt_bfgs_step_finish(self.dtype_code, self.B, self.n, self.state.data_ptr(), self.res.data_ptr(),
                                                     self._stream()), "b200opt_bfgs_step_finish")

    # carry layout (csrc/generic_step.cu): [H (n*n) | grad (n) | newton (n) | f | alpha | first | terminate | num_accepted | ...]
    def terminate_flags(self):
        n = self.n
        return self.state[:, n * n + 2 * n + 3] != 0

    @property
    def first_step(self):
        n = self.n
        return self._sq(self.state[:, n * n + 2 * n + 2] != 0)

    @property
    def search_state(self):
        n = self.n
        return self._sq(self.state[:, n * n + 2 * n + 1])

    @property
    def num_accepted_steps(self):
        n = self.n
        return self._sq(self.state[:, n * n + 2 * n + 4].to(self.t.int32))

    @property
    def y_eval(self):
        return self._sq(self.y_eval_buf)

    @property
    def f_info(self):
        n, B = self.n, self.B
        H = self.state[:, : n * n].reshape(B, n, n)
        f, g = self.state[:, n * n + 2 * n], self.state[:, n * n : n * n + n]
        cls = FunctionInfo.EvalGradHessianInv if self.use_inverse else FunctionInfo.EvalGradHessian
        return cls(self._sq(f), self._sq(g), self._sq(H))


# ------------------------------------------------------------------------------------------ protocol methods
def _split_aux(fn):
    """The protocol's `fn` returns `(out, aux)` (interactive.ipynb: "all functions are assumed to return auxiliary
    information"); a plain `fn(y, args) -> out` and a RegisteredModel (callable on torch tensors) are accepted too."""
    def first(y, a):
        out = fn(y, a)
        return out[0] if isinstance(out, tuple) else out

    def aux_of(y, a):
        out = fn(y, a)
        return out[1] if isinstance(out, tuple) else None

    return first, aux_of


def protocol_init(solver, session_cls, fn, y, args, options):
    first, aux_of = _split_aux(fn)
    opts = dict(options or {})
    in_dims = opts.pop("args_in_dims", None)
    sess = session_cls(first, solver, y, args, opts or None, args_in_dims=in_dims, max_steps=None)
    if sess.empty:
        raise ValueError("init: the batch is empty - there is nothing to step through (the entry points return an empty Solution)")
    sess._aux_of = aux_of
    sess._in_dims = in_dims
    return sess


def protocol_step(state, args):
    """One step for every problem that has not stopped.  Returns `(new_y, state, aux)`: the accepted iterate, the (same,
    updated) state and the auxiliary output of `fn` at it."""
    state.step()
    new_y = state._sq(state.y)
    return new_y, state, _aux_at(state, new_y, args)


def _aux_at(state, y, args):
    """The auxiliary output of `fn` at `y` (`None` when the function has none), batched like `y`."""
    t = _torch()
    aux_of = getattr(state, "_aux_of", None)
    if aux_of is None or getattr(state, "_aux_is_none", False):
        return None
    with t.no_grad():
        if y.ndim == 1:
            return aux_of(y, args)
        saw_none = []

        def tensor_aux(v, a):  # vmap wants tensors out: a function without auxiliary output returns a placeholder
            out = aux_of(v, a)
            if out is None:
                saw_none.append(True)
                return v.new_zeros(())
            return out

        if state._in_dims is None:
            out = t.func.vmap(lambda v: tensor_aux(v, args))(y)
        else:
            out = t.func.vmap(tensor_aux, in_dims=(0, state._in_dims))(y, args)
        if saw_none:
            state._aux_is_none = True  # decided once: later steps skip the evaluation
            return None
        return out
