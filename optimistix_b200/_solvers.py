"""Solver-class surface of the reference, kept name-for-name so call sites do not change.

Constructor fields follow the reference:
  LevenbergMarquardt  optimistix/_solver/levenberg_marquardt.py:281-294
  GaussNewton         optimistix/_solver/gauss_newton.py:382-395
  Newton / Chord      optimistix/_solver/newton_chord.py:59-65
  BFGS                optimistix/_solver/quasi_newton.py:443-458
  ClassicalTrustRegion optimistix/_solver/trust_region.py:125-128, BacktrackingArmijo backtracking.py:29-31,
  LearningRate learning_rate.py:18, DampedNewtonDescent / NewtonDescent (descents).

These are plain immutable descriptions: the arithmetic happens in the CUDA kernels, which implement the
DEFAULT search/descent combination of each solver.  Non-default searches/descents raise at call time.
"""
from __future__ import annotations

import dataclasses
from typing import Any, Callable


# ---- norms (optimistix/_misc.py:41-45) ------------------------------------------------------------------------
def _norm(name):
    def deco(fn):
        fn.norm_name = name
        return fn

    return deco


@_norm("max")
def max_norm(x):
    """Largest absolute value (NaN-propagating), like `optimistix.max_norm`."""
    import torch

    x = torch.as_tensor(x)
    return x.abs().max() if x.numel() else x.new_zeros(())


@_norm("two")
def two_norm(x):
    import torch

    return torch.as_tensor(x).square().sum().sqrt()


@_norm("rms")
def rms_norm(x):
    import torch

    return torch.as_tensor(x).square().mean().sqrt()


# ---- lineax stand-ins (only used as tokens; see DESIGN.md "linear solves") -----------------------------------------
@dataclasses.dataclass(frozen=True)
class QR:
    """`lx.QR()`: selects the kernels' Householder-QR descent (B200OPT_LINSOLVE_QR) when given to LevenbergMarquardt / GaussNewton."""


@dataclasses.dataclass(frozen=True)
class NormalCholesky:
    """Cholesky of J^T J + lambda I with J never materialised (BASELINE north_star; `lx.NormalCholesky` is the nearest lineax
    name): the default linear solve of the CUDA least-squares kernels.  The reference's default for LevenbergMarquardt is
    `lx.QR()` (levenberg_marquardt.py:286): pass `linear_solver=QR()` to get the reference's own operations."""


@dataclasses.dataclass(frozen=True)
class Cholesky:
    pass


@dataclasses.dataclass(frozen=True)
class LU:
    pass


@dataclasses.dataclass(frozen=True)
class AutoLinearSolver:
    well_posed: Any = None


# ---- searches and descents ------------------------------------------------------------------------------------
class AbstractSearch:
    pass


class AbstractDescent:
    pass


@dataclasses.dataclass(frozen=True)
class ClassicalTrustRegion(AbstractSearch):
    high_cutoff: float = 0.99
    low_cutoff: float = 0.01
    high_constant: float = 3.5
    low_constant: float = 0.25


@dataclasses.dataclass(frozen=True)
class BacktrackingArmijo(AbstractSearch):
    decrease_factor: float = 0.5
    slope: float = 0.1
    step_init: float = 1.0


@dataclasses.dataclass(frozen=True)
class LearningRate(AbstractSearch):
    learning_rate: float = 1.0


@dataclasses.dataclass(frozen=True)
class DampedNewtonDescent(AbstractDescent):
    linear_solver: Any = AutoLinearSolver(well_posed=None)


@dataclasses.dataclass(frozen=True)
class NewtonDescent(AbstractDescent):
    norm: Callable | None = None
    linear_solver: Any = AutoLinearSolver(well_posed=None)


# ---- solver hierarchy -----------------------------------------------------------------------------------------
def _session(kind):
    from . import _protocol

    return {"lstsq": _protocol.LsqSession, "root": _protocol.RootSession, "min": _protocol.MinSession}[kind]


class _ProtocolBase:
    """`init / step / terminate / postprocess` with the reference's signatures (optimistix/_iterate.py:33-184), backed by the
    per-iteration CUDA step kernels: see `_protocol.py`."""

    _kind = ""

    def init(self, fn, y, args, options, f_struct=None, aux_struct=None, tags=frozenset()):
        from . import _protocol

        return _protocol.protocol_init(self, _session(self._kind), fn, y, args, options)

    def step(self, fn, y, args, options, state, tags=frozenset()):
        from . import _protocol

        return _protocol.protocol_step(state, args)

    def terminate(self, fn, y, args, options, state, tags=frozenset()):
        return state.done, state.result

    def postprocess(self, fn, y, aux, args, options, state, tags=frozenset(), result=None):
        return y, aux, {}


class AbstractIterativeSolver:
    pass


class AbstractLeastSquaresSolver(AbstractIterativeSolver):
    pass


class AbstractRootFinder(AbstractIterativeSolver):
    pass


class AbstractMinimiser(AbstractIterativeSolver):
    pass


class AbstractFixedPointSolver(AbstractIterativeSolver):
    pass


class AbstractGaussNewton(AbstractLeastSquaresSolver, _ProtocolBase):
    _kind = "lstsq"


class AbstractQuasiNewton(AbstractMinimiser, _ProtocolBase):
    _kind = "min"


class AbstractBFGS(AbstractQuasiNewton):
    pass


class AbstractDFP(AbstractQuasiNewton):
    pass


def _default_verbose(**kwargs):
    """`verbose=True` (optimistix/_misc.py:385-393): one line per step, every field."""
    pieces = []
    for name, value in kwargs.values():
        if hasattr(value, "detach"):
            value = value.detach().cpu()
            value = value.item() if value.ndim == 0 else value.tolist()
        pieces.append(f"{name}: {value}")
    if pieces:
        print(", ".join(pieces))


def _check_verbose(verbose):
    """`verbose: bool | Callable[..., None]` as in the reference (gauss_newton.py:300-312, quasi_newton.py:273-283).  Returns
    None (silent) or the callable that receives `name=(label, value)` keyword arguments once per step.  A verbose solve runs
    through the per-iteration step kernels (the host sees every step); the persistent kernels never print."""
    if verbose in (False, None):
        return None
    if verbose is True:
        return _default_verbose
    if callable(verbose):
        return verbose
    raise TypeError("verbose must be a bool or a callable `**kwargs -> None`")


class LevenbergMarquardt(AbstractGaussNewton):
    def __init__(self, rtol: float, atol: float, norm: Callable = max_norm, linear_solver: Any = NormalCholesky(), verbose=False):
        self.verbose = _check_verbose(verbose)
        self.rtol, self.atol, self.norm = rtol, atol, norm
        self.descent = DampedNewtonDescent(linear_solver=linear_solver)
        self.search = ClassicalTrustRegion()


class GaussNewton(AbstractGaussNewton):
    def __init__(self, rtol: float, atol: float, norm: Callable = max_norm,
                 linear_solver: Any = NormalCholesky(), verbose=False):
        self.verbose = _check_verbose(verbose)
        self.rtol, self.atol, self.norm = rtol, atol, norm
        self.descent = NewtonDescent(linear_solver=linear_solver)
        self.search = LearningRate(1.0)


class _AbstractNewtonChord(AbstractRootFinder, _ProtocolBase):
    _is_newton: bool
    _kind = "root"

    def __init__(self, rtol: float, atol: float, norm: Callable = max_norm, kappa: float = 1e-2,
                 linear_solver: Any = AutoLinearSolver(well_posed=None), cauchy_termination: bool = True):
        self.rtol, self.atol, self.norm, self.kappa = rtol, atol, norm, kappa
        self.linear_solver, self.cauchy_termination = linear_solver, cauchy_termination


class Newton(_AbstractNewtonChord):
    _is_newton = True


class Chord(_AbstractNewtonChord):
    _is_newton = False


class BFGS(AbstractBFGS):
    """quasi_newton.py:443-458.  use_inverse=True (default): registered families run in the persistent kernel (H^-1 in
    registers); Python callables and use_inverse=False (Hessian approximation + Cholesky solve) run in the generic
    per-iteration step kernel."""
    _update = "bfgs"

    def __init__(self, rtol: float, atol: float, norm: Callable = max_norm, use_inverse: bool = True, verbose=False):
        self.verbose = _check_verbose(verbose)
        self.rtol, self.atol, self.norm, self.use_inverse = rtol, atol, norm, use_inverse
        self.descent = NewtonDescent(linear_solver=Cholesky())
        self.search = BacktrackingArmijo()


class DFP(AbstractDFP):
    """Davidon-Fletcher-Powell (quasi_newton.py:576-633): same search and descent as BFGS, the DFP update
    (:516-547).  Runs in the generic per-iteration step kernel."""
    _update = "dfp"

    def __init__(self, rtol: float, atol: float, norm: Callable = max_norm, use_inverse: bool = True, verbose=False):
        self.verbose = _check_verbose(verbose)
        self.rtol, self.atol, self.norm, self.use_inverse = rtol, atol, norm, use_inverse
        self.descent = NewtonDescent(linear_solver=Cholesky())
        self.search = BacktrackingArmijo()


# ---- adjoints (names kept; differentiating through the CUDA solve is SURVEY 8f.2) ---------------------------------
class AbstractAdjoint:
    pass


@dataclasses.dataclass(frozen=True)
class ImplicitAdjoint(AbstractAdjoint):
    linear_solver: Any = AutoLinearSolver(well_posed=None)


@dataclasses.dataclass(frozen=True)
class RecursiveCheckpointAdjoint(AbstractAdjoint):
    checkpoints: int | None = None
