"""Generate tests/golden/*.npz from the reference-faithful oracle (QR linear solves).

The reference itself cannot run in this image (no jax), so these vectors are outputs of the NumPy
restatement in oracle/solvers.py, which is pinned against the reference's known answers by
tests/test_oracle_known_answers.py.  Inputs are regenerated from seeds (optimistix_b200/synthetic.py),
only outputs are stored.  Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from optimistix_b200 import synthetic  # noqa: E402
from oracle import models as M  # noqa: E402
from oracle import solvers as S  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
GOLDEN_BATCH = 512


def save(name, out, **extra):
    np.savez_compressed(os.path.join(OUT, name + ".npz"), value=out["value"], num_steps=out["num_steps"], result=out["result"], **extra)
    print(name, "steps", np.bincount(out["num_steps"])[:40], "results", np.bincount(out["result"], minlength=8))


def main():
    w = synthetic.exp_decay_batch(GOLDEN_BATCH)
    a = dict(t=w["t"], d=w["d"])
    for ls in ("qr", "gram"):
        save(f"lm_exp_decay_f64_{ls}", S.levenberg_marquardt(M.exp_decay, w["y0"], a, rtol=1e-8, atol=1e-8, linsolve=ls))
    save("gn_exp_decay_f64_qr", S.gauss_newton(M.exp_decay, w["y0"], a, rtol=1e-8, atol=1e-8, linsolve="qr"))
    w = synthetic.gauss2_batch(GOLDEN_BATCH)
    a = dict(x=w["x"], d=w["d"])
    for ls in ("qr", "gram"):
        save(f"lm_gauss2_f32_{ls}", S.levenberg_marquardt(M.gauss2, w["y0"], a, rtol=1e-4, atol=1e-4, linsolve=ls))
    w = synthetic.implicit_euler_batch(GOLDEN_BATCH)
    save("newton_implicit_euler_f64", S.newton_chord(M.implicit_euler, w["y0"], w, rtol=1e-8, atol=1e-8))
    save("chord_implicit_euler_f64", S.newton_chord(M.implicit_euler, w["y0"], w, rtol=1e-8, atol=1e-8, chord=True))
    w = synthetic.rosenbrock_batch(64)
    save("bfgs_rosenbrock_f64", S.bfgs(M.rosenbrock, w["y0"], None, rtol=1e-8, atol=1e-8, max_steps=4096))


if __name__ == "__main__":
    main()
