"""Host->device copy rate from pinned memory on this box (context for the e2e number: fit is H2D-bound)."""
import json
import time

import torch

n = 1 << 30
src = torch.empty(n, dtype=torch.uint8, pin_memory=True)
dst = torch.empty(n, dtype=torch.uint8, device="cuda")
out = {}
for label, chunks in (("1GiB_one_copy", 1), ("8x128MiB", 8)):
    best = 0.0
    for _ in range(3):
        torch.cuda.synchronize()
        t = time.perf_counter()
        step = n // chunks
        for c in range(chunks):
            dst[c * step:(c + 1) * step].copy_(src[c * step:(c + 1) * step], non_blocking=True)
        torch.cuda.synchronize()
        best = max(best, n / (time.perf_counter() - t) / 1e9)
    out[label + "_GBps"] = round(best, 2)
print(json.dumps(out))
