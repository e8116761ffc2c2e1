"""FunctionInfo containers (optimistix_b200/_function_info.py; reference optimistix/_search.py:44-183). CPU."""
import numpy as np
import torch

import optimistix_b200 as optx

FI = optx.FunctionInfo


def test_variants_and_names():
    for name in ("Eval", "EvalGrad", "EvalGradHessian", "EvalGradHessianInv", "Residual", "ResidualJac"):
        cls = getattr(FI, name)
        assert issubclass(cls, FI) and cls.__qualname__ == f"FunctionInfo.{name}"  # _search.py:186-197


def test_minimisation_variants():
    g = torch.tensor([[1.0, 2.0, 3.0], [0.5, 0.0, -1.0]], dtype=torch.float64)
    y = torch.tensor([[1.0, 1.0, 1.0], [2.0, 3.0, 4.0]], dtype=torch.float64)
    f = torch.tensor([3.0, 4.0], dtype=torch.float64)
    for info in (FI.EvalGrad(f, g), FI.EvalGradHessian(f, g, torch.eye(3).expand(2, 3, 3)), FI.EvalGradHessianInv(f, g, torch.eye(3).expand(2, 3, 3))):
        assert torch.equal(info.as_min(), f)
        assert torch.allclose(info.compute_grad_dot(y), torch.tensor([6.0, -3.0], dtype=torch.float64))
    assert torch.equal(FI.Eval(f).as_min(), f)


def test_least_squares_variants_against_numpy():
    rng = np.random.default_rng(0)
    J = rng.standard_normal((5, 7, 3))
    r = rng.standard_normal((5, 7))
    y = rng.standard_normal((5, 3))
    info = FI.ResidualJac(torch.from_numpy(r), torch.from_numpy(J))
    assert np.allclose(info.as_min().numpy(), 0.5 * (r * r).sum(1))                       # 0.5 * sum_squares, _search.py:141-142
    assert np.allclose(info.compute_grad().numpy(), np.einsum("bmn,bm->bn", J, r))      # J^T r, :153-155
    assert np.allclose(info.compute_grad_dot(torch.from_numpy(y)).numpy(), np.einsum("bm,bmn,bn->b", r, J, y))  # r^T J y, :157-179
    assert np.allclose(FI.Residual(torch.from_numpy(r)).as_min().numpy(), 0.5 * (r * r).sum(1))
    # the Gram-carried form of the CUDA kernels answers the same questions without J
    gram = info.to_gram()
    assert np.allclose(gram.as_min().numpy(), info.as_min().numpy())
    assert np.allclose(gram.compute_grad().numpy(), info.compute_grad().numpy())
    assert np.allclose(gram.compute_grad_dot(torch.from_numpy(y)).numpy(), info.compute_grad_dot(torch.from_numpy(y)).numpy())
    d = 0.1 * y
    pred = 0.5 * (((np.einsum("bmn,bn->bm", J, d) + r) ** 2).sum(1) - (r * r).sum(1))    # trust_region.py:182-186
    assert np.allclose(gram.predicted_reduction(torch.from_numpy(d)).numpy(), pred)


def test_unbatched():
    J = torch.tensor([[1.0, 0.0], [0.0, 2.0], [1.0, 1.0]])
    r = torch.tensor([1.0, -1.0, 2.0])
    info = FI.ResidualJac(r, J)
    assert torch.allclose(info.compute_grad(), torch.tensor([3.0, 0.0]))
    assert float(info.as_min()) == 3.0
