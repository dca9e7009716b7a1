"""Where the LK step on region-of-interest pyramids spends its time (config 3 batch, one GPU)."""
import sys, ctypes as C, numpy as np, torch
sys.path.insert(0, '.')
import bench
from accurate_aprilgroup_tracking_b200 import synth
from accurate_aprilgroup_tracking_b200.context import AgtContext
CAM = bench.CAM
ctx = AgtContext(0, CAM.mtx, None)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
traj = np.array([synth.trajectory(3000 + i, 2) for i in range(B)])
pa, pb = ctx.alloc_pyramid(B, CAM.width, CAM.height, 4), ctx.alloc_pyramid(B, CAM.width, CAM.height, 4)
for b0 in range(0, B, 512):
    nb = min(512, B - b0)
    ctx.render(pa, traj[b0:b0 + nb, 0], np.arange(nb) + b0, offset=b0, batch=nb)
    ctx.render(pb, traj[b0:b0 + nb, 1], np.arange(nb) + b0 + 1, offset=b0, batch=nb)
ctx.build_pyramid(pa); ctx.build_pyramid(pb)
obj = synth.object_points()
pts = torch.as_tensor(np.stack([synth.project(obj, traj[i, 0], CAM) for i in range(B)]).astype(np.float32), device=ctx.tdev)
t = torch
out = t.empty_like(pts); st = t.empty((B, 48), dtype=t.uint8, device=ctx.tdev); err = t.empty((B, 48), dtype=t.float32, device=ctx.tdev)
left = t.empty((B, 48), dtype=t.uint8, device=ctx.tdev); redo = t.empty(B, dtype=t.uint8, device=ctx.tdev)
rects = ctx.lk_rects(pb, pts, None, 32)
r = rects.cpu().numpy()
print("mean rect %.0f x %.0f px, area %.3f of the frame" % ((r[:, 2] - r[:, 0]).mean(), (r[:, 3] - r[:, 1]).mean(),
      ((r[:, 2] - r[:, 0]) * (r[:, 3] - r[:, 1])).mean() / (CAM.width * CAM.height)))
p = ctx._p
args = (p(pts), p(out), p(st), p(err), None)
def timed(label, fn, n=10):
    for _ in range(3): fn()
    e0, e1 = t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); t.cuda.synchronize()
    print(f"{label:44s} {e0.elapsed_time(e1) / n:8.3f} ms", flush=True)
ctx._use_current_stream()
timed("lk_rects", lambda: ctx.lk_rects(pb, pts, None, 32))
timed("build_pyramid (complete)", lambda: ctx.build_pyramid(pb))
timed("build_pyramid_roi", lambda: ctx.build_pyramid_roi(pb, rects))
ctx.build_pyramid(pb)
timed("lk (complete pyramids)", lambda: ctx.lk(pa, pb, pts))
timed("lk_roi kernel, rects, no mask", lambda: ctx._check(ctx.lib.agt_lk_roi(ctx.h, C.byref(pa.desc), C.byref(pb.desc), *args, None, p(rects), 4, None, p(left), B, 48)))
timed("any_flag", lambda: ctx._check(ctx.lib.agt_any_flag(ctx.h, p(left), 48, p(redo), B)))
print("frames flagged:", int(redo.sum()), "corners flagged:", int(left.sum()))
timed("build_pyramid_masked", lambda: ctx.build_pyramid_masked(pb, redo))
timed("lk masked (complete)", lambda: ctx._check(ctx.lib.agt_lk_roi(ctx.h, C.byref(pa.desc), C.byref(pb.desc), *args, None, None, 0, p(redo), None, B, 48)))
timed("lk_roi (whole step)", lambda: ctx.lk_roi(pa, pb, pts))
