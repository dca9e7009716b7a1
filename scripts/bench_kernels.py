"""Micro-benchmarks of the individual kernels (run on the GPU box): pyramid levels, LK, PnP."""
import sys, numpy as np
sys.path.insert(0, '.')
import torch, ctypes as C
from accurate_aprilgroup_tracking_b200 import synth
from accurate_aprilgroup_tracking_b200.context import AgtContext
cam = synth.CAMERA_1080P
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
ctx = AgtContext(0, cam.mtx, None); ctx.set_synthetic_model()
rng = np.random.default_rng(1)
traj = np.array([synth.trajectory(100 + i, 2) for i in range(B)])
pa = ctx.alloc_pyramid(B, cam.width, cam.height, 4); pb = ctx.alloc_pyramid(B, cam.width, cam.height, 4)
for b0 in range(0, B, 256):
    nb = min(256, B - b0)
    ctx.render(pa, traj[b0:b0+nb, 0], np.arange(nb) + b0, offset=b0, batch=nb)
    ctx.render(pb, traj[b0:b0+nb, 1], np.arange(nb) + b0 + 7, offset=b0, batch=nb)
def timeit(fn, n=10):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
lib, h = ctx.lib, ctx.h
ctx._use_current_stream()
for l in range(3):
    d = pa.desc
    def f(l=l):
        ctx._check(lib.agt_pyr_down(h, C.c_void_p(d.data[l]), d.width[l], d.height[l], d.pitch[l], d.frame_stride[l],
                                    C.c_void_p(d.data[l+1]), d.pitch[l+1], d.frame_stride[l+1], B))
    ms = timeit(f)
    byt = B * (d.width[l]*d.height[l] + d.width[l+1]*d.height[l+1])
    print(f"pyr_down L{l}->L{l+1}: {ms:.3f} ms  {byt/ms/1e6:.0f} GB/s ({byt/ms/1e6/6550.1:.2f} of measured HBM peak)")
ms = timeit(lambda: ctx.build_pyramid(pa))
print(f"build_pyramid (3 levels) {ms:.3f} ms -> {2754000*B/ms/1e6:.0f} GB/s algorithmic ({2754000*B/ms/1e6/6550.1:.2f})")
ctx.build_pyramid(pb)
obj = synth.object_points()
pts = np.stack([synth.project(obj, traj[i, 0], cam) for i in range(B)]).astype(np.float32)
dpts = torch.as_tensor(pts, device=ctx.tdev)
ms = timeit(lambda: ctx.lk(pa, pb, dpts))
out, st, err = ctx.lk(pa, pb, dpts)
print(f"lk: {ms:.3f} ms for {B}x48 corners -> {B*48/ms*1e3:.3e} corners/s, {5008*B*48/ms/1e6:.0f} GB/s algorithmic ({5008*B*48/ms/1e6/6550.1:.3f}); tracked {float(st.float().mean()):.3f}")
valid = np.zeros((B, 48), np.uint8)
for i in range(B):
    for k in synth.visible_tags(traj[i, 1]): valid[i, 4*k:4*k+4] = 1
dvalid = torch.as_tensor(valid, device=ctx.tdev)
ms = timeit(lambda: ctx.pnp(obj.astype(np.float32), out, dvalid))
print(f"pnp (DLT init): {ms:.3f} ms for {B} frames -> {B/ms*1e3:.3e} poses/s")
pose, ok, e, it = ctx.pnp(obj.astype(np.float32), out, dvalid)
ms = timeit(lambda: ctx.pnp(obj.astype(np.float32), out, dvalid, pose, torch.ones(B, dtype=torch.uint8, device=ctx.tdev)))
print(f"pnp (guess): {ms:.3f} ms for {B} frames -> {B/ms*1e3:.3e} poses/s; ok {float(ok.float().mean()):.3f}")
init = pose.reshape(B, 1, 6)
ms = timeit(lambda: ctx.refine(pa if False else pb, init, 1), 5)
res = ctx.refine(pb, init, 1)
ev = res["evals"].double(); nv = res["n_valid"].double()
print(f"refine: {ms:.3f} ms for {B} poses -> {B/ms*1e3:.3e} poses/s; mean evals {float(ev.mean()):.2f}; algorithmic {float((36*ev*nv).sum())/ms/1e6:.0f} GB/s")
# frame ingest: BGR -> gray, and undistort + crop + BGR -> gray (row N2)
import cv2
nb = min(B, 256)
bgr = torch.randint(0, 256, (nb, cam.height, cam.width, 3), dtype=torch.uint8, device=ctx.tdev)
pg = ctx.alloc_pyramid(nb, cam.width, cam.height, 1)
ms = timeit(lambda: ctx.ingest_bgr(pg, bgr))
byt = nb * cam.width * cam.height * 4
print(f"bgr_to_gray: {ms:.3f} ms for {nb} 1080p frames -> {byt/ms/1e6:.0f} GB/s ({byt/ms/1e6/6550.1:.2f} of measured HBM peak)")
dist = np.array([[-0.28, 0.11, 0.0007, -0.0004, -0.02]])
new_mtx, roi = cv2.getOptimalNewCameraMatrix(cam.mtx, dist, (cam.width, cam.height), 1, (cam.width, cam.height))
ctxd = AgtContext(0, cam.mtx, dist)
ctxd.set_undistort(new_mtx, cam.width, cam.height, roi)
pu = ctxd.alloc_pyramid(nb, roi[2], roi[3], 1)
ms = timeit(lambda: ctxd.ingest_undistort(pu, bgr))
byt = nb * (cam.width * cam.height * 3 + roi[2] * roi[3])
print(f"undistort_to_gray: {ms:.3f} ms for {nb} 1080p BGR frames (roi {roi[2]}x{roi[3]}) -> {nb/ms*1e3:.3e} frames/s, {byt/ms/1e6:.0f} GB/s ({byt/ms/1e6/6550.1:.2f} of measured HBM peak)")
