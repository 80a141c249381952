#!/usr/bin/env python
"""GPU probe: pinned H2D of a 4096^2 u8 image, 1-D copy vs 2-D copy into a pitched destination."""
import torch, time
n = 4096
src = torch.empty((n, n), dtype=torch.uint8, pin_memory=True)
dst1 = torch.empty((n, n), dtype=torch.uint8, device="cuda")
dst2 = torch.empty((n, n + 128), dtype=torch.uint8, device="cuda")
def t(f, reps=20):
    for _ in range(3): f()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): f()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps
a = t(lambda: dst1.copy_(src, non_blocking=True))
b = t(lambda: dst2[:, :n].copy_(src, non_blocking=True))
print(f"1-D {a*1e3:.3f} ms = {n*n/a/1e9:.1f} GB/s;  2-D pitched {b*1e3:.3f} ms = {n*n/b/1e9:.1f} GB/s")
