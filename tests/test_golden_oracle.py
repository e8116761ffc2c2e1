"""The oracle reproduces the committed golden vectors (regression pin of the restatement itself). CPU."""
import os

import numpy as np
import pytest

from optimistix_b200 import synthetic
from oracle import models as M
from oracle import solvers as S

G = os.path.join(os.path.dirname(__file__), "golden")
NB = 64  # prefix of the golden batch re-solved here (batched solve == per-lane solve, so a prefix is exact)


def _load(name):
    return np.load(os.path.join(G, name + ".npz"))


@pytest.mark.parametrize("ls", ("qr", "gram"))
def test_lm_exp_decay_golden(ls):
    w = synthetic.exp_decay_batch(512)
    out = S.levenberg_marquardt(M.exp_decay, w["y0"][:NB], dict(t=w["t"], d=w["d"][:NB]), rtol=1e-8, atol=1e-8, linsolve=ls)
    g = _load(f"lm_exp_decay_f64_{ls}")
    assert np.array_equal(out["num_steps"], g["num_steps"][:NB])
    assert np.array_equal(out["result"], g["result"][:NB])
    assert np.allclose(out["value"], g["value"][:NB], rtol=1e-12, atol=0)


def test_lm_gauss2_golden():
    w = synthetic.gauss2_batch(512)
    out = S.levenberg_marquardt(M.gauss2, w["y0"][:NB], dict(x=w["x"], d=w["d"][:NB]), rtol=1e-4, atol=1e-4, linsolve="gram")
    g = _load("lm_gauss2_f32_gram")
    assert np.mean(out["num_steps"] == g["num_steps"][:NB]) >= 0.95  # fp32 BLAS reductions may differ by batch shape
    assert np.allclose(out["value"], g["value"][:NB], rtol=1e-3, atol=1e-3)


def test_newton_golden():
    w = synthetic.implicit_euler_batch(512)
    a = dict(A=w["A"], B=w["B"], y_prev=w["y_prev"][:NB], dt=w["dt"][:NB])
    out = S.newton_chord(M.implicit_euler, w["y0"][:NB], a, rtol=1e-8, atol=1e-8)
    g = _load("newton_implicit_euler_f64")
    assert np.array_equal(out["num_steps"], g["num_steps"][:NB])
    assert np.allclose(out["value"], g["value"][:NB], rtol=1e-12, atol=0)


def test_bfgs_golden():
    w = synthetic.rosenbrock_batch(64)
    out = S.bfgs(M.rosenbrock, w["y0"][:8], None, rtol=1e-8, atol=1e-8, max_steps=4096)
    g = _load("bfgs_rosenbrock_f64")
    assert np.array_equal(out["result"], g["result"][:8])
    assert np.allclose(out["value"], g["value"][:8], rtol=1e-6, atol=1e-8)


def test_c_newton_and_bfgs_oracles_match_the_numpy_oracle():
    """oracle/c/nb_oracle.c (the multi-core CPU baseline of bench.py for configs[2] / configs[3]) against oracle/solvers.py."""
    from optimistix_b200 import synthetic
    from oracle import c_oracle, models as M, solvers as S

    w = synthetic.implicit_euler_batch(512)
    a = dict(A=w["A"], B=w["B"], y_prev=w["y_prev"], dt=w["dt"])
    for chord in (False, True):
        ref = S.newton_chord(M.implicit_euler, w["y0"], a, rtol=1e-8, atol=1e-8, chord=chord)
        c = c_oracle.newton_implicit_euler(w["y0"], w["A"], w["B"], w["y_prev"], w["dt"], rtol=1e-8, atol=1e-8, chord=chord)
        assert np.array_equal(c["num_steps"], ref["num_steps"]) and np.array_equal(c["result"], ref["result"])
        assert np.allclose(c["value"], ref["value"], rtol=1e-13, atol=1e-15)
    w = synthetic.rosenbrock_batch(16)
    for k in (1, 3, 10):
        ref = S.bfgs(M.rosenbrock, w["y0"], None, rtol=1e-8, atol=1e-8, max_steps=k)
        c = c_oracle.bfgs_rosenbrock(w["y0"], rtol=1e-8, atol=1e-8, max_steps=k)
        assert np.array_equal(c["num_steps"], ref["num_steps"]) and np.array_equal(c["result"], ref["result"])
        assert np.allclose(c["value"], ref["value"], rtol=1e-12, atol=1e-14)
    ref = S.bfgs(M.rosenbrock, w["y0"], None, rtol=1e-8, atol=1e-8, max_steps=4096)
    c = c_oracle.bfgs_rosenbrock(w["y0"], rtol=1e-8, atol=1e-8, max_steps=4096)
    assert np.array_equal(c["result"], ref["result"])
    fc, _ = M.rosenbrock(c["value"])
    fr, _ = M.rosenbrock(ref["value"])
    assert np.allclose(fc, fr, atol=1e-8)
