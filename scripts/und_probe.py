"""One timed (or profiled) run of the undistort ingest on 256 x 1080p BGR frames."""
import sys, numpy as np, torch, cv2
sys.path.insert(0, '.')
if len(sys.argv) > 1:
    from pathlib import Path
    from accurate_aprilgroup_tracking_b200 import _lib
    _lib.LIB_PATH = Path(sys.argv[1]).resolve()
from accurate_aprilgroup_tracking_b200 import synth
from accurate_aprilgroup_tracking_b200.context import AgtContext
cam = synth.CAMERA_1080P
nb = 256
dist = np.array([[-0.28, 0.11, 0.0007, -0.0004, -0.02]])
new_mtx, roi = cv2.getOptimalNewCameraMatrix(cam.mtx, dist, (cam.width, cam.height), 1, (cam.width, cam.height))
ctx = AgtContext(0, cam.mtx, dist)
ctx.set_undistort(new_mtx, cam.width, cam.height, roi)
bgr = torch.randint(0, 256, (nb, cam.height, cam.width, 3), dtype=torch.uint8, device=ctx.tdev)
pu = ctx.alloc_pyramid(nb, roi[2], roi[3], 1)
for _ in range(3): ctx.ingest_undistort(pu, bgr)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): ctx.ingest_undistort(pu, bgr)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
byt = nb * (cam.width * cam.height * 3 + roi[2] * roi[3])
import hashlib
h = hashlib.sha256(pu.frames.cpu().numpy().tobytes()).hexdigest()[:12]
print(f"{sys.argv[1] if len(sys.argv) > 1 else 'libagt.so'} [{h}] undistort_to_gray: {ms:.3f} ms for {nb} 1080p BGR frames (roi {roi[2]}x{roi[3]}) -> {nb/ms*1e3:.3e} frames/s, {byt/ms/1e6:.0f} GB/s ({byt/ms/1e6/6550.1:.2f} of measured HBM peak)")
