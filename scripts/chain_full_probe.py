"""Full-frame pyramid: three per-level launches against the single-launch chain kernel fed with whole-frame rectangles."""
import sys, numpy as np, torch
sys.path.insert(0, '.')
import bench
from accurate_aprilgroup_tracking_b200.context import AgtContext
CAM = bench.CAM
ctx = AgtContext(0, CAM.mtx, None)
B = 2048
pb = ctx.alloc_pyramid(B, CAM.width, CAM.height, 4)
pb.levels[0].copy_(torch.randint(0, 256, pb.levels[0].shape, dtype=torch.uint8, device=ctx.tdev))
ref = ctx.alloc_pyramid(B, CAM.width, CAM.height, 4)
ref.levels[0].copy_(pb.levels[0])
ctx.build_pyramid(ref)
rects = torch.tensor([[0, 0, CAM.width, CAM.height]], dtype=torch.int32, device=ctx.tdev).repeat(B, 1).contiguous()
def timed(label, fn, n=10):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(f"{label:40s} {ms:8.3f} ms  -> {B * 2754000 / ms / 1e6:7.0f} GB/s algorithmic ({B * 2754000 / ms / 1e6 / 6550.1:.2f})", flush=True)
timed("per-level launches (agt_build_pyramid)", lambda: ctx.build_pyramid(pb))
timed("one launch, CTA per frame (chain)", lambda: ctx.build_pyramid_roi(pb, rects))
print("equal:", all(bool(torch.equal(pb.levels[l], ref.levels[l])) for l in (1, 2, 3)))
