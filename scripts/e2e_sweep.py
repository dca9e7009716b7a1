import os, sys, time, numpy as np
sys.path.insert(0, '.')
import torch
from accurate_aprilgroup_tracking_b200 import synth
from accurate_aprilgroup_tracking_b200.context import AgtContext
from accurate_aprilgroup_tracking_b200.cv_compat import HostContext
import bench
CAM = synth.CAMERA_1080P
B = 4096
ctx = AgtContext(0, CAM.mtx, None)
truth, init = bench.make_poses(B, 2000)
pyr = ctx.alloc_pyramid(B, CAM.width, CAM.height, 1)
for b0 in range(0, B, 512):
    ctx.render(pyr, truth[b0:b0+512], np.arange(b0, b0+512) + 2000, offset=b0, batch=512)
host_frames = torch.empty((B, CAM.height, CAM.width), dtype=torch.uint8, pin_memory=True)
host_frames.copy_(pyr.frames); torch.cuda.synchronize()
hf = host_frames.numpy()
del pyr
h = HostContext(0)
s, tg, n, c = synth.surface_model(); h.set_model(s, tg, n, c, synth.model_pitch())
for chunk in (1024, 512, 256, 128):
    for ctas in (32, 64, 128, 256):
        os.environ["AGT_E2E_CHUNK"] = str(chunk); os.environ["AGT_GATHER_CTAS"] = str(ctas)
        for _ in range(2): h.refine_poses(hf, init, CAM.mtx)
        t0 = time.perf_counter()
        for _ in range(4): h.refine_poses(hf, init, CAM.mtx)
        dt = (time.perf_counter() - t0) / 4
        print(f"chunk {chunk:5d} gather CTAs {ctas:4d}: {dt*1e3:7.2f} ms  {B/dt:9.0f} poses/s", flush=True)
