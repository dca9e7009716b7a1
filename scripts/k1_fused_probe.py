"""K1 in one pass (pyr_fused_kernel) against the per-level launches: time per batch and bit-equality of levels 1-3 (run on the GPU
box).  The switch is read at agt_create: AGT_K1_FUSED=1 selects the one-pass kernel (default: per-level launches).
    python scripts/k1_fused_probe.py [frames] [vga]"""
import os, sys, numpy as np
sys.path.insert(0, '.')
import torch
from accurate_aprilgroup_tracking_b200 import synth
from accurate_aprilgroup_tracking_b200.context import AgtContext
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
cam = synth.CAMERA_VGA if len(sys.argv) > 2 and sys.argv[2] == "vga" else synth.CAMERA_1080P
def timeit(fn, n=10):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
os.environ["AGT_K1_FUSED"] = "0"
ref = AgtContext(0, cam.mtx, None)
os.environ["AGT_K1_FUSED"] = "1"
fus = AgtContext(0, cam.mtx, None)
pa = ref.alloc_pyramid(B, cam.width, cam.height, 4)
g = torch.Generator(device="cuda"); g.manual_seed(1)
# smooth random texture (blurred noise) so that rounding in the 5-tap sums is exercised everywhere
for b0 in range(0, B, 256):
    nb = min(256, B - b0)
    x = torch.randint(0, 256, (nb, 1, cam.height, cam.width), device="cuda", generator=g, dtype=torch.uint8).float()
    x = torch.nn.functional.avg_pool2d(x, 3, 1, 1)
    pa.frames[b0:b0 + nb].copy_(x[:, 0].round().to(torch.uint8))
bytes_alg = sum(pa.desc.width[l] * pa.desc.height[l] for l in range(4)) * B
t_ref = timeit(lambda: ref.build_pyramid(pa))
want = [pa.levels[l].clone() for l in range(1, 4)]
for l in range(1, 4): pa.levels[l].zero_()
t_fus = timeit(lambda: fus.build_pyramid(pa))
same = [bool(torch.equal(pa.levels[l], want[l - 1])) for l in range(1, 4)]
print(f"{cam.width}x{cam.height} x {B}: per-level {t_ref:.3f} ms ({bytes_alg / t_ref / 1e6 / 6550.1:.3f} of peak), one pass {t_fus:.3f} ms "
      f"({bytes_alg / t_fus / 1e6 / 6550.1:.3f} of peak), levels identical {same}")
assert all(same)
