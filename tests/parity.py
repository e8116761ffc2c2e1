"""Shared parity helpers: compare a CUDA solve with an oracle / golden solve of the same inputs."""
from __future__ import annotations

import numpy as np

from optimistix_b200 import RESULTS


def to_np(x):
    return x.detach().cpu().numpy() if hasattr(x, "detach") else np.asarray(x)


def compare(sol, ref, rtol, atol, label=""):
    """Return a dict of parity statistics between a Solution (CUDA) and an oracle result dict."""
    val, steps, res = to_np(sol.value), to_np(sol.stats["num_steps"]), to_np(sol.result)
    rv, rs, rr = ref["value"], ref["num_steps"], ref["result"]
    same_steps = steps == rs
    # difference in units of the solver's own termination scale (atol + rtol*|y|)
    scaled = np.max(np.abs(val - rv) / (atol + rtol * np.abs(rv)), axis=1)
    rel = np.max(np.abs(val - rv) / np.maximum(np.abs(rv), 1e-300), axis=1)
    both_ok = (res == RESULTS.successful) & (rr == RESULTS.successful)  # diverged / max-steps iterates are not comparable
    stats = dict(
        n=len(steps),
        result_match=float(np.mean(res == rr)),
        steps_match=float(np.mean(same_steps)),
        mean_steps=(float(steps.mean()), float(rs.mean())),
        scaled_same_max=float(scaled[same_steps].max()) if same_steps.any() else 0.0,
        scaled_p50=float(np.percentile(scaled, 50)),
        scaled_p99=float(np.percentile(scaled, 99)),
        scaled_max=float(scaled.max()),
        rel_same_p99=float(np.percentile(rel[same_steps], 99)) if same_steps.any() else 0.0,
        rel_same_max=float(rel[same_steps].max()) if same_steps.any() else 0.0,
        within_tol=float(np.mean(scaled <= 1.0)),
        both_ok=float(np.mean(both_ok)),
        within_tol_ok=float(np.mean(scaled[both_ok] <= 1.0)) if both_ok.any() else 1.0,
        within_2tol_ok=float(np.mean(scaled[both_ok] <= 2.0)) if both_ok.any() else 1.0,
    )
    print(f"[parity {label}] " + " ".join(f"{k}={v}" for k, v in stats.items()))
    return stats


def result_names(res):
    return [RESULTS.name(c) for c in to_np(res)]
