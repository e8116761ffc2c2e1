"""Bitwise A/B check of two builds / variants of the packed LM kernel on the gauss2 family (fp32, m = 256).

    python tools/lm_ab_check.py save /tmp/ref.npz                        # default kernel
    B200OPT_LSQ_LEAN=1 python tools/lm_ab_check.py save /tmp/lean.npz    # lsq_kernel_p2l.cuh (or B200OPT_LIB=... for a build)
    python tools/lm_ab_check.py cmp /tmp/ref.npz /tmp/lean.npz           # variants that only rearrange the work must be bit-identical

(The switch is read once per process, hence the separate runs.)
"""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import optimistix_b200 as optx  # noqa: E402
from optimistix_b200 import synthetic  # noqa: E402


def cases():
    out = {}
    c = lambda a: torch.from_numpy(a).cuda()
    for name, batch, solver, max_steps, poison in [
        ("lm_20000", 20000, optx.LevenbergMarquardt(1e-4, 1e-4), 256, False),
        ("gn_5000", 5000, optx.GaussNewton(1e-4, 1e-4), 256, False),
        ("lm_1", 1, optx.LevenbergMarquardt(1e-4, 1e-4), 256, False),
        ("lm_3", 3, optx.LevenbergMarquardt(1e-4, 1e-4), 256, False),
        ("lm_1001_ms1", 1001, optx.LevenbergMarquardt(1e-4, 1e-4), 1, False),
        ("lm_1001_ms2", 1001, optx.LevenbergMarquardt(1e-4, 1e-4), 2, False),
        ("lm_1001_ms5", 1001, optx.LevenbergMarquardt(1e-4, 1e-4), 5, False),
        ("lm_777_nan", 777, optx.LevenbergMarquardt(1e-4, 1e-4), 256, True),
        ("lm_9000_tight", 9000, optx.LevenbergMarquardt(1e-7, 1e-7), 64, False),
    ]:
        w = synthetic.gauss2_batch(batch, m=256, seed=11)
        if poison:
            w["y0"][5, 2] = np.nan
            w["y0"][300, 0] = np.inf
            w["d"][17, 100] = np.nan
            w["y0"][40, 2] = 0.0  # sigma = 0: singular / non-finite path
        sol = optx.least_squares(optx.models.gauss2, solver, c(w["y0"]), (c(w["x"]), c(w["d"])), throw=False, max_steps=max_steps)
        torch.cuda.synchronize()
        out[name + ".y"] = sol.value.cpu().numpy()
        out[name + ".steps"] = sol.stats["num_steps"].cpu().numpy()
        out[name + ".res"] = sol.result.cpu().numpy()
        out[name + ".delta"] = sol.stats["step_size"].cpu().numpy()
        out[name + ".nacc"] = sol.stats["num_accepted_steps"].cpu().numpy()
    return out


if sys.argv[1] == "save":
    np.savez(sys.argv[2], **cases())
    print("saved", sys.argv[2])
else:
    a, b = np.load(sys.argv[2]), np.load(sys.argv[3])
    bad = 0
    for k in a.files:
        same = np.array_equal(a[k], b[k], equal_nan=True)
        if not same:
            bad += 1
            diff = np.flatnonzero(~((a[k] == b[k]) | (np.isnan(a[k].astype(np.float64)) & np.isnan(b[k].astype(np.float64)))).reshape(len(a[k]), -1).all(1))
            print("DIFF", k, "rows", diff[:8], "of", len(a[k]), "n_diff", diff.size)
            r = diff[0]
            print("   ref", a[k][r], "\n   new", b[k][r])
        else:
            print("same", k, a[k].shape)
    print("BIT-IDENTICAL" if bad == 0 else f"{bad} arrays differ")
