"""LK flow against cv2.calcOpticalFlowPyrLK on a 64-pair subset of the config-3 batch (1080p, 48 corners per pair)."""
import sys, numpy as np, torch
sys.path.insert(0, '.')
import bench
from accurate_aprilgroup_tracking_b200 import synth
from accurate_aprilgroup_tracking_b200.context import AgtContext
from oracle import lk_oracle
CAM = bench.CAM
ctx = AgtContext(0, CAM.mtx, None)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
traj = np.array([synth.trajectory(3000 + i, 2) for i in range(B)])
pa, pb = ctx.alloc_pyramid(B, CAM.width, CAM.height, 4), ctx.alloc_pyramid(B, CAM.width, CAM.height, 4)
ctx.render(pa, traj[:, 0], np.arange(B)); ctx.render(pb, traj[:, 1], np.arange(B) + 1)
ctx.build_pyramid(pa); ctx.build_pyramid(pb)
obj = synth.object_points()
pts = np.stack([synth.project(obj, traj[i, 0], CAM) for i in range(B)]).astype(np.float32)
out, st, err = [t.cpu().numpy() for t in ctx.lk(pa, pb, pts)]
fa, fb = pa.frames.cpu().numpy(), pb.frames.cpu().numpy()
diffs, status_mismatch, n = [], 0, 0
vis_d, occ_d = [], []
for i in range(B):
    ro, rs, re = lk_oracle.lk_cv(fa[i], fb[i], pts[i])
    status_mismatch += int((st[i] != rs).sum())
    m = (rs == 1) & (st[i] == 1)
    dd = np.abs(out[i] - ro).max(axis=1)
    diffs.append(dd[m])
    n += int(m.sum())
    vis = np.zeros(48, bool)
    for k in synth.visible_tags(traj[i, 0]):
        vis[4 * k:4 * k + 4] = True
    vis_d.append(dd[m & vis]); occ_d.append(dd[m & ~vis])
    for j in np.nonzero(m & (dd > 0.005))[0]:
        print(f"  pair {i} corner {j}: diff {dd[j]:.4f} px, visible tag {bool(vis[j])}, err gpu {err[i][j]:.3f} cv {re[j]:.3f}, flow {np.linalg.norm(ro[j] - pts[i][j]):.2f} px")
d = np.concatenate(diffs)
vd, od = np.concatenate(vis_d), np.concatenate(occ_d)
print(f"corners of visible tags: {vd.size}, max {vd.max():.4e}, above 0.01: {int((vd > 0.01).sum())}; corners of hidden tags (tracking whatever covers them): {od.size}, max {od.max():.4e}, above 0.01: {int((od > 0.01).sum())}")
print(f"{B} pairs, {n} tracked corners: status mismatches {status_mismatch}; flow difference max {d.max():.4e} px, p99.9 {np.quantile(d, 0.999):.4e}, "
      f"p99 {np.quantile(d, 0.99):.4e}, median {np.median(d):.4e}; corners above 0.01 px: {int((d > 0.01).sum())}, above 0.005 px: {int((d > 0.005).sum())}")
