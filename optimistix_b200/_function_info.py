"""`FunctionInfo`: what a solver knows about the objective at a point.

Mirrors the variants and method names of the reference (optimistix/_search.py:44-183: `Eval`, `EvalGrad`,
`EvalGradHessian`, `EvalGradHessianInv`, `Residual`, `ResidualJac`; `as_min`, `compute_grad`, `compute_grad_dot`) for code
that inspects a solver state.  Two differences, by construction:

* operators are dense tensors, not lineax operators (`jac` is `[..., m, n]`, `hessian` / `hessian_inv` are `[..., n, n]`);
  every method broadcasts over leading batch axes, the explicit counterpart of the reference's `jax.vmap`;
* the CUDA kernels never materialise `J`.  What they carry for an accepted point is `(r^T r, G = J^T J, g = J^T r)`
  (csrc/lsq_kernel.cuh), which is the variant `ResidualGram` here: it answers the same questions the searches ask of
  `ResidualJac` (`as_min`, `compute_grad`, `compute_grad_dot`) plus the predicted reduction
  `0.5 (|J d + r|^2 - |r|^2) = d.g + 0.5 d.G d` of trust_region.py:182-186.

Plain containers on the host side of the boundary: nothing here runs on the hot path.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, ClassVar


def _dot(a, b):
    return (a * b).sum(-1)


class FunctionInfo:
    """Enumeration-ish base class: `FunctionInfo.{Eval, EvalGrad, EvalGradHessian, EvalGradHessianInv, Residual,
    ResidualJac, ResidualGram}`."""

    Eval: ClassVar[type]
    EvalGrad: ClassVar[type]
    EvalGradHessian: ClassVar[type]
    EvalGradHessianInv: ClassVar[type]
    Residual: ClassVar[type]
    ResidualJac: ClassVar[type]
    ResidualGram: ClassVar[type]

    def as_min(self):
        """f(y) for a minimisation problem, 0.5 * sum(residual^2) for a least-squares problem."""
        raise NotImplementedError


@dataclass(frozen=True)
class Eval(FunctionInfo):
    f: Any

    def as_min(self):
        return self.f


@dataclass(frozen=True)
class EvalGrad(FunctionInfo):
    f: Any
    grad: Any

    def as_min(self):
        return self.f

    def compute_grad_dot(self, y):
        return _dot(self.grad, y)


@dataclass(frozen=True)
class EvalGradHessian(FunctionInfo):
    f: Any
    grad: Any
    hessian: Any  # [..., n, n]

    def as_min(self):
        return self.f

    def compute_grad_dot(self, y):
        return _dot(self.grad, y)


@dataclass(frozen=True)
class EvalGradHessianInv(FunctionInfo):
    f: Any
    grad: Any
    hessian_inv: Any  # [..., n, n]

    def as_min(self):
        return self.f

    def compute_grad_dot(self, y):
        return _dot(self.grad, y)


@dataclass(frozen=True)
class Residual(FunctionInfo):
    residual: Any  # [..., m]

    def as_min(self):
        return 0.5 * (self.residual * self.residual).sum(-1)


@dataclass(frozen=True)
class ResidualJac(FunctionInfo):
    residual: Any  # [..., m]
    jac: Any       # [..., m, n]

    def as_min(self):
        return 0.5 * (self.residual * self.residual).sum(-1)

    def compute_grad(self):
        return (self.jac * self.residual[..., :, None]).sum(-2)  # J^T r

    def compute_grad_dot(self, y):
        # residual^T (J y): the reduction order the reference chooses (forward mode), _search.py:157-179
        return _dot((self.jac * y[..., None, :]).sum(-1), self.residual)

    def to_gram(self):
        """The form the CUDA kernels carry."""
        J, r = self.jac, self.residual
        return ResidualGram(sum_squares=(r * r).sum(-1), gram=J.transpose(-1, -2) @ J, grad=self.compute_grad())


@dataclass(frozen=True)
class ResidualGram(FunctionInfo):
    sum_squares: Any  # [...]        r^T r
    gram: Any         # [..., n, n]  J^T J
    grad: Any         # [..., n]     J^T r

    def as_min(self):
        return 0.5 * self.sum_squares

    def compute_grad(self):
        return self.grad

    def compute_grad_dot(self, y):
        return _dot(self.grad, y)

    def predicted_reduction(self, y_diff):
        """0.5 (|J d + r|^2 - |r|^2) without J (trust_region.py:182-186)."""
        return _dot(self.grad, y_diff) + 0.5 * _dot(y_diff, (self.gram * y_diff[..., None, :]).sum(-1))


for _c in (Eval, EvalGrad, EvalGradHessian, EvalGradHessianInv, Residual, ResidualJac, ResidualGram):
    _c.__qualname__ = "FunctionInfo." + _c.__name__
    setattr(FunctionInfo, _c.__name__, _c)
