"""Where the LM iteration-count bias of the round-1 kernels came from, shown on the CPU with the oracle alone.

The CUDA kernels carry (r, G = J^T J, g = J^T r, r.r) instead of the reference's (r, J).  The predicted reduction of
ClassicalTrustRegion, 0.5 (|J d + r|^2 - |r|^2) (trust_region.py:182-186), is then available as d.g + 0.5 d.G d - the
same number in exact arithmetic, but NOT in floating point: the reference's value is quantised to the spacing of |r|^2
(exactly 0 for a step too small to move |r|^2), and the accept test `df <= 0.01 pred` near convergence depends on it
(a step with df == 0 is accepted by the reference, rejected when pred is an exactly evaluated tiny negative number).
The exact form (round 1) takes +1.5-3 % more iterations than the reference arithmetic; the kernels' form (Gram identity,
then rounded and subtracted like the reference: csrc/common.cuh predicted_reduction) stays within the spread between
faithful restatements (measured -0.2 ... -0.7 %; the QR and Cholesky solves alone differ by 0.2-0.4 %, NumPy and C
restatements of the same mode by as much: the reference's own value of `pred` for a step below the resolution of |r|^2 is
the rounding noise of two m-term sums, i.e. defined by the summation order of whoever evaluates it).  The GPU tests assert the
same bound against the QR oracle (tests/test_lm_gpu.py)."""
import numpy as np
import pytest

from optimistix_b200 import synthetic
from oracle import c_oracle
from oracle import models as M
from oracle import solvers as S

CASES = {
    "exp_decay": (synthetic.exp_decay_batch, "t", 1e-8),
    "gauss2": (synthetic.gauss2_batch, "x", 1e-4),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_pred_mode_bias_c_oracle(name):
    make, xk, tol = CASES[name]
    w = make(8192)
    run = lambda mode: c_oracle.levenberg_marquardt(name, w["y0"], w[xk], w["d"], rtol=tol, atol=tol, linsolve=mode)  # noqa: E731
    ref = run("qr")
    stats = {}
    for mode in ("gram", "kernel", "kernel_exact_pred"):
        o = run(mode)
        assert (o["result"] == ref["result"]).all()
        bias = o["num_steps"].mean() / ref["num_steps"].mean() - 1.0
        match = float((o["num_steps"] == ref["num_steps"]).mean())
        scaled = np.abs(o["value"] - ref["value"]) / (tol + tol * np.abs(ref["value"]))
        stats[mode] = (bias, match)
        print(f"[pred modes {name}] {mode}: mean steps {o['num_steps'].mean():.4f} vs QR {ref['num_steps'].mean():.4f} "
              f"bias {100 * bias:+.2f}% count match {match:.3f} within tol {(scaled.max(1) <= 1).mean():.4f}")
        assert (scaled.max(1) <= 1.0).mean() >= 0.99
    assert stats["kernel_exact_pred"][0] > 0.008          # the exactly evaluated form: biased upwards
    assert abs(stats["kernel"][0]) <= 0.008               # the kernels' arithmetic: within the spread of faithful restatements
    assert abs(stats["gram"][0]) <= 0.006
    assert abs(stats["kernel"][0]) < stats["kernel_exact_pred"][0]
    assert stats["kernel"][1] > stats["kernel_exact_pred"][1]


def test_pred_modes_numpy_oracle_agree_with_c():
    """The NumPy oracle's `pred` modes are the same arithmetic as the C restatement's: same bias picture on exp_decay."""
    w = synthetic.exp_decay_batch(1024)
    args = dict(t=w["t"], d=w["d"])
    ref = S.levenberg_marquardt(M.exp_decay, w["y0"], args, rtol=1e-8, atol=1e-8, linsolve="qr", pred="residual")
    q = S.levenberg_marquardt(M.exp_decay, w["y0"], args, rtol=1e-8, atol=1e-8, linsolve="qr", pred="gram")
    e = S.levenberg_marquardt(M.exp_decay, w["y0"], args, rtol=1e-8, atol=1e-8, linsolve="qr", pred="gram_exact")
    bq = q["num_steps"].mean() / ref["num_steps"].mean() - 1.0
    be = e["num_steps"].mean() / ref["num_steps"].mean() - 1.0
    mq = (q["num_steps"] == ref["num_steps"]).mean()
    me = (e["num_steps"] == ref["num_steps"]).mean()
    print(f"[pred modes numpy exp_decay] gram {100 * bq:+.2f}% (count match {mq:.3f}), gram_exact {100 * be:+.2f}% (count match {me:.3f})")
    assert abs(bq) <= 0.008 and be > 0.01 and mq > me and mq >= 0.97
    # C restatement, kernel-arithmetic mode, against the NumPy one on the same problems: same result codes, values to tolerance
    k_np = S.levenberg_marquardt(M.exp_decay, w["y0"], args, rtol=1e-8, atol=1e-8, linsolve="gram", pred="gram")
    k_c = c_oracle.levenberg_marquardt("exp_decay", w["y0"], w["t"], w["d"], rtol=1e-8, atol=1e-8, linsolve="kernel")
    assert (k_np["result"] == k_c["result"]).all()
    scaled = np.abs(k_np["value"] - k_c["value"]) / (1e-8 + 1e-8 * np.abs(k_c["value"]))
    assert scaled.max() <= 1.0
    assert abs(k_np["num_steps"].mean() / k_c["num_steps"].mean() - 1.0) <= 0.01


@pytest.mark.skipif(not c_oracle.fma_available(), reason="needs an x86 CPU with FMA3 for the contracted build of the C oracle")
@pytest.mark.parametrize("name", sorted(CASES))
def test_iteration_counts_move_with_the_rounding_alone(name):
    """The yardstick for final-iteration-count parity.  liboracle_fma.so is liboracle.so compiled with a*b+c -> fma (what nvcc
    does to the CUDA kernels by default): the same algorithm, the same operation order, only the roundings differ.  Already
    that changes the final count on 20-40 % of the problems and lowers the mean by a few tenths of a per cent - every problem
    of these workloads ends on accept / reject decisions inside the rounding noise of |r|^2, and fewer roundings mean more
    exact ties `df == 0`, which the reference's test accepts.  The CUDA kernels against the oracle sit at exactly this level
    (count match 0.58-0.64, bias -0.2 ... -1.1 %: profiles/lm_parity_r2*.json), so problem-by-problem count equality with ANY
    independent implementation - XLA's included - is not a property the reference algorithm has on these workloads."""
    make, xk, tol = CASES[name]
    w = make(8192)
    for mode in ("qr", "kernel"):
        a = c_oracle.levenberg_marquardt(name, w["y0"], w[xk], w["d"], rtol=tol, atol=tol, linsolve=mode)
        b = c_oracle.levenberg_marquardt(name, w["y0"], w[xk], w["d"], rtol=tol, atol=tol, linsolve=mode, fma=True)
        match = float((a["num_steps"] == b["num_steps"]).mean())
        bias = b["num_steps"].mean() / a["num_steps"].mean() - 1.0
        scaled = np.abs(a["value"] - b["value"]) / (tol + tol * np.abs(a["value"]))
        print(f"[fma yardstick {name} {mode}] count match {match:.3f}, mean-step bias of the contracted build {100 * bias:+.2f}%, "
              f"values within tol {(scaled.max(1) <= 1).mean():.4f}")
        assert (a["result"] == b["result"]).mean() >= 0.999
        assert (scaled.max(1) <= 1.0).mean() >= 0.99
        assert match <= 0.9          # far from identical although nothing but the rounding changed
        assert abs(bias) <= 0.012


@pytest.mark.skipif(not c_oracle.fma_available(), reason="needs an x86 CPU with FMA3 for the contracted build of the C oracle")
def test_bfgs_moves_with_the_rounding_alone():
    """The same yardstick for BFGS on the chained Rosenbrock function (configs[3]): plain against contracted build of the same
    C code - about a third of the counts equal, ~95 % of the minimisers within one tolerance unit, all within ten."""
    w = synthetic.rosenbrock_batch(2048)
    a = c_oracle.bfgs_rosenbrock(w["y0"], rtol=1e-8, atol=1e-8, max_steps=4096)
    b = c_oracle.bfgs_rosenbrock(w["y0"], rtol=1e-8, atol=1e-8, max_steps=4096, fma=True)
    scaled = np.max(np.abs(a["value"] - b["value"]) / (1e-8 + 1e-8 * np.abs(a["value"])), axis=1)
    match = float((a["num_steps"] == b["num_steps"]).mean())
    print(f"[fma yardstick bfgs] count match {match:.3f}, within 1 tol {np.mean(scaled <= 1):.4f}, within 10 tol {np.mean(scaled <= 10):.4f}")
    assert (a["result"] == b["result"]).all()
    assert match <= 0.6 and 0.9 <= np.mean(scaled <= 1.0) <= 0.99 and np.mean(scaled <= 10.0) >= 0.999
