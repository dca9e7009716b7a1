"""Time agt_lk of one build of the library (argv[1] = path of a libagt.so variant) on the config-3 batch and print a checksum
of its outputs, so that variants built with different staging parameters can be compared in one GPU call."""
import sys, hashlib, numpy as np, torch
from pathlib import Path
sys.path.insert(0, '.')
from accurate_aprilgroup_tracking_b200 import _lib
_lib.LIB_PATH = Path(sys.argv[1]).resolve()
import bench
from accurate_aprilgroup_tracking_b200 import synth
from accurate_aprilgroup_tracking_b200.context import AgtContext
CAM = bench.CAM
ctx = AgtContext(0, CAM.mtx, None)
B = 2048
traj = np.array([synth.trajectory(3000 + i, 2) for i in range(B)])
pa, pb = ctx.alloc_pyramid(B, CAM.width, CAM.height, 4), ctx.alloc_pyramid(B, CAM.width, CAM.height, 4)
for b0 in range(0, B, 512):
    ctx.render(pa, traj[b0:b0 + 512, 0], np.arange(512) + b0, offset=b0, batch=512)
    ctx.render(pb, traj[b0:b0 + 512, 1], np.arange(512) + b0 + 1, offset=b0, batch=512)
ctx.build_pyramid(pa); ctx.build_pyramid(pb)
obj = synth.object_points()
pts = torch.as_tensor(np.stack([synth.project(obj, traj[i, 0], CAM) for i in range(B)]).astype(np.float32), device=ctx.tdev)
for _ in range(3): out = ctx.lk(pa, pb, pts)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): out = ctx.lk(pa, pb, pts)
e1.record(); torch.cuda.synchronize()
h = hashlib.sha256(b"".join(o.cpu().numpy().tobytes() for o in out)).hexdigest()[:16]
print(f"{Path(sys.argv[1]).name:28s} {e0.elapsed_time(e1) / 20:7.3f} ms per {B * 48} corners   outputs {h}", flush=True)
