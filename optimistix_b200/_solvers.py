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
    pass


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


class AbstractGaussNewton(AbstractLeastSquaresSolver):
    pass


class AbstractQuasiNewton(AbstractMinimiser):
    pass


class AbstractBFGS(AbstractQuasiNewton):
    pass


class AbstractDFP(AbstractQuasiNewton):
    pass


def _check_verbose(verbose):
    if verbose not in (False, None):
        raise NotImplementedError("verbose=... printing is not available from inside the persistent CUDA kernel")


class LevenbergMarquardt(AbstractGaussNewton):
    def __init__(self, rtol: float, atol: float, norm: Callable = max_norm, linear_solver: Any = QR(), verbose=False):
        _check_verbose(verbose)
        self.rtol, self.atol, self.norm = rtol, atol, norm
        self.descent = DampedNewtonDescent(linear_solver=linear_solver)
        self.search = ClassicalTrustRegion()
        self.verbose = verbose


class GaussNewton(AbstractGaussNewton):
    def __init__(self, rtol: float, atol: float, norm: Callable = max_norm,
                 linear_solver: Any = AutoLinearSolver(well_posed=None), verbose=False):
        _check_verbose(verbose)
        self.rtol, self.atol, self.norm = rtol, atol, norm
        self.descent = NewtonDescent(linear_solver=linear_solver)
        self.search = LearningRate(1.0)
        self.verbose = verbose


class _AbstractNewtonChord(AbstractRootFinder):
    _is_newton: bool

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
        _check_verbose(verbose)
        self.rtol, self.atol, self.norm, self.use_inverse = rtol, atol, norm, use_inverse
        self.descent = NewtonDescent(linear_solver=Cholesky())
        self.search = BacktrackingArmijo()
        self.verbose = verbose


class DFP(AbstractDFP):
    """Davidon-Fletcher-Powell (quasi_newton.py:576-633): same search and descent as BFGS, the DFP update
    (:516-547).  Runs in the generic per-iteration step kernel."""
    _update = "dfp"

    def __init__(self, rtol: float, atol: float, norm: Callable = max_norm, use_inverse: bool = True, verbose=False):
        _check_verbose(verbose)
        self.rtol, self.atol, self.norm, self.use_inverse = rtol, atol, norm, use_inverse
        self.descent = NewtonDescent(linear_solver=Cholesky())
        self.search = BacktrackingArmijo()
        self.verbose = verbose


# ---- adjoints (names kept; differentiating through the CUDA solve is SURVEY 8f.2) ---------------------------------
class AbstractAdjoint:
    pass


@dataclasses.dataclass(frozen=True)
class ImplicitAdjoint(AbstractAdjoint):
    linear_solver: Any = AutoLinearSolver(well_posed=None)


@dataclasses.dataclass(frozen=True)
class RecursiveCheckpointAdjoint(AbstractAdjoint):
    checkpoints: int | None = None
