"""Host-to-device copy bandwidth of this box: one large pinned copy on the copy engine (what the ROI gather competes with)."""
import time, torch
n = 400 << 20
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
for chunk in (n, n // 4, n // 16, 1 << 20):
    torch.cuda.synchronize()
    for _ in range(2):
        for o in range(0, n, chunk):
            d[o:o + chunk].copy_(h[o:o + chunk], non_blocking=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        for o in range(0, n, chunk):
            d[o:o + chunk].copy_(h[o:o + chunk], non_blocking=True)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 5
    print(f"pinned H2D {n >> 20} MiB in pieces of {chunk >> 10} KiB: {dt * 1e3:.2f} ms = {n / dt / 1e9:.1f} GB/s")
