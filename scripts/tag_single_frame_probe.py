"""The detector as the drop-in uses it (detect_pose.py:368-371: one gray frame in host memory, detections back on the host):
wall time of apriltag_gpu.Detector.detect against cv2.aruco on the same frame, and the device part alone."""
import sys, time, numpy as np, torch
sys.path.insert(0, '.')
import cv2
from accurate_aprilgroup_tracking_b200 import synth, apriltag_gpu, cv_compat
from accurate_aprilgroup_tracking_b200.context import AgtContext
from oracle import tag_oracle
cv2.setNumThreads(4)                                          # the reference's detector runs with nthreads=4 (detect_pose.py:88)
for cam in (synth.CAMERA_VGA, synth.CAMERA_1080P):
    ctx = AgtContext(0, cam.mtx, None)
    pose = synth.trajectory(700, 1)[0]
    pyr = ctx.alloc_pyramid(1, cam.width, cam.height, 1)
    ctx.render(pyr, pose[None], np.arange(1))
    gray = pyr.frames[0].cpu().numpy()
    det = apriltag_gpu.Detector()
    for _ in range(5): found = det.detect(gray)
    t0 = time.perf_counter()
    for _ in range(50): found = det.detect(gray)
    host_ms = (time.perf_counter() - t0) / 50 * 1e3
    for _ in range(3): ctx.detect_tags(pyr)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(50): ctx.detect_tags(pyr)
    e1.record(); torch.cuda.synchronize()
    dev_ms = e0.elapsed_time(e1) / 50
    for _ in range(3): ar = tag_oracle.detect_cv(gray)
    t0 = time.perf_counter()
    for _ in range(20): ar = tag_oracle.detect_cv(gray)
    cv_ms = (time.perf_counter() - t0) / 20 * 1e3
    print(f"{cam.width}x{cam.height}: Detector.detect (host frame in, detections out) {host_ms:.3f} ms, device part {dev_ms:.3f} ms, "
          f"cv2.aruco (4 threads) {cv_ms:.2f} ms; tags {sorted(d.tag_id for d in found)} / aruco {sorted(i for i, _ in ar)}")
    ctx.close()
