"""Host<->device copy bandwidth on this box (pinned vs pageable), to put the e2e number in context."""
import time, json, torch
n = 1 << 28  # 1 GiB of float32
res = {}
for kind in ("pinned", "pageable"):
    h = torch.empty(n, dtype=torch.float32, pin_memory=(kind == "pinned"))
    h.fill_(1.0)
    d = torch.empty(n, dtype=torch.float32, device="cuda")
    for direction in ("h2d", "d2h"):
        best = 1e9
        for _ in range(3):
            torch.cuda.synchronize(); t = time.perf_counter()
            (d.copy_(h, non_blocking=True) if direction == "h2d" else h.copy_(d, non_blocking=True))
            torch.cuda.synchronize(); best = min(best, time.perf_counter() - t)
        res[f"{kind}_{direction}_GBps"] = round(4 * n / best / 1e9, 2)
print(json.dumps(res))
