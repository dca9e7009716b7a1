"""A few ROI pyramid builds on the config-3 batch (for ncu)."""
import sys, numpy as np, torch
sys.path.insert(0, '.')
import bench
from accurate_aprilgroup_tracking_b200 import synth
from accurate_aprilgroup_tracking_b200.context import AgtContext
CAM = bench.CAM
ctx = AgtContext(0, CAM.mtx, None)
B = 1024
traj = np.array([synth.trajectory(3000 + i, 2) for i in range(B)])
pb = ctx.alloc_pyramid(B, CAM.width, CAM.height, 4)
for b0 in range(0, B, 512):
    ctx.render(pb, traj[b0:b0 + 512, 1], np.arange(512) + b0 + 1, offset=b0, batch=512)
obj = synth.object_points()
pts = torch.as_tensor(np.stack([synth.project(obj, traj[i, 0], CAM) for i in range(B)]).astype(np.float32), device=ctx.tdev)
rects = ctx.lk_rects(pb, pts, None, 32)
for _ in range(3): ctx.build_pyramid_roi(pb, rects)
torch.cuda.synchronize()
print("ok")
