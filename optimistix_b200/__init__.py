"""optimistix_b200 - B200-native (sm_100a) implementation of the Optimistix data-parallel hot path.

Drop-in surface for the hot path of patrick-kidger/optimistix 0.1.0 (optimistix/__init__.py):
`least_squares`, `minimise`, `root_find`, `fixed_point`, the solver classes `LevenbergMarquardt`,
`GaussNewton`, `Newton`, `Chord`, `BFGS`, `DFP`, `compat.minimize`, the norms, `FunctionInfo`, `RESULTS` and `Solution`; plus `models`, the
residual-model plugin registry.  Everything numerical runs in hand-written CUDA kernels behind the
C ABI of include/b200opt.h (optimistix_b200/libb200opt.so); there is no CPU fallback.
"""
from . import _models as models
from . import compat
from ._entry import fixed_point, least_squares, minimise, root_find
from ._function_info import FunctionInfo
from ._solution import RESULTS, Solution
from ._solvers import (
    BFGS,
    DFP,
    AbstractAdjoint,
    AbstractBFGS,
    AbstractDFP,
    AbstractDescent,
    AbstractFixedPointSolver,
    AbstractGaussNewton,
    AbstractIterativeSolver,
    AbstractLeastSquaresSolver,
    AbstractMinimiser,
    AbstractQuasiNewton,
    AbstractRootFinder,
    AbstractSearch,
    AutoLinearSolver,
    BacktrackingArmijo,
    Cholesky,
    Chord,
    ClassicalTrustRegion,
    DampedNewtonDescent,
    GaussNewton,
    ImplicitAdjoint,
    LearningRate,
    LevenbergMarquardt,
    LU,
    Newton,
    NewtonDescent,
    NormalCholesky,
    QR,
    RecursiveCheckpointAdjoint,
    max_norm,
    rms_norm,
    two_norm,
)

__version__ = "0.1.0"
