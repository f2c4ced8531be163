"""Raw pinned host<->device copy bandwidth of the box (context for bench.py's e2e figure)."""
import time
import torch
n = 1 << 30
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
for name, a, b in (("h2d", d, h), ("d2h", h, d)):
    for _ in range(2):
        a.copy_(b, non_blocking=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        a.copy_(b, non_blocking=True)
    torch.cuda.synchronize()
    print(name, f"{5 * n / (time.perf_counter() - t0) / 1e9:.1f} GB/s")
# both directions at once on two streams (what cfg.async_pull of the krmdup path relies on)
h2 = torch.empty(n, dtype=torch.uint8).pin_memory()
d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    with torch.cuda.stream(s1):
        d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2):
        h2.copy_(d2, non_blocking=True)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print("h2d + d2h concurrently", f"{5 * n / dt / 1e9:.1f} GB/s each way, {10 * n / dt / 1e9:.1f} GB/s total")
