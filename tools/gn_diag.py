import sys, os, numpy as np, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import optimistix_b200 as optx
from optimistix_b200 import synthetic
from oracle import solvers as S, models as M
for m in (64, 256):
    w = synthetic.gauss2_batch(1024, m=m, seed=5)
    c = lambda a: torch.from_numpy(a).cuda()
    sol = optx.least_squares(optx.models.gauss2, optx.GaussNewton(1e-4, 1e-4), c(w["y0"]), (c(w["x"]), c(w["d"])), throw=False)
    y = sol.value.cpu().numpy(); res = sol.result.cpu().numpy() if hasattr(sol.result, "cpu") else np.asarray(sol.result)
    r32 = S.gauss_newton(M.gauss2, w["y0"], dict(x=w["x"], d=w["d"]), rtol=1e-4, atol=1e-4, linsolve="gram")
    r64 = S.gauss_newton(M.gauss2, w["y0"].astype(np.float64), dict(x=w["x"].astype(np.float64), d=w["d"].astype(np.float64)), rtol=1e-4, atol=1e-4, linsolve="gram")
    def close(a, b): return np.all(np.abs(a - b) <= 2 * (1e-4 + 1e-4 * np.abs(b)), axis=-1)
    ok32 = r32["result"] == 0; ok64 = r64["result"] == 0
    print(m, os.environ.get("B200OPT_LSQ_SCALAR"), "gpu~o32", close(y, r32["value"])[ok32].mean(), "gpu~o64", close(y, r64["value"])[ok64].mean(),
          "o32~o64", close(r32["value"], r64["value"])[ok64 & ok32].mean(), "ok", ok32.mean(), ok64.mean(), (np.asarray(res) == 0).mean())
