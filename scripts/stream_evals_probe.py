"""How many evaluations does the slowest refinement of a 64-stream frame-step run (the cluster K4 launch lasts as long as it)?"""
import sys, numpy as np, torch
sys.path.insert(0, '.')
import bench
from accurate_aprilgroup_tracking_b200 import synth
from accurate_aprilgroup_tracking_b200.batched import BatchedPoseDetector, pack_detections
from accurate_aprilgroup_tracking_b200.context import AgtContext
CAM = bench.CAM
ctx = AgtContext(0, CAM.mtx, None); ctx.set_synthetic_model()
S, F = 64, 32
trajs = [synth.trajectory(5000 + s, F) for s in range(S)]
rngs = [np.random.default_rng(5000 + s) for s in range(S)]
bank = ctx.alloc_pyramid(S, CAM.width, CAM.height, 1)
bpd = BatchedPoseDetector(ctx, S, CAM.width, CAM.height, synth.object_points())
ev, it, nv = [], [], []
for f in range(F):
    ctx.render(bank, np.array([trajs[i][f] for i in range(S)]), np.array([1000 * s + f for s in range(S)]))
    rows = []
    for i in range(S):
        d = synth.detections(trajs[i][f], CAM, rngs[i])
        if (f + 3 * i) % 17 == 16: d = d[:1]
        rows.append(d)
    a, b, c = bpd.pack(rows)
    out = bpd.step(a, b, c, frames=bank.frames)
    r = out["refine"]
    st = r["status"].cpu().numpy().ravel()
    e = r["evals"].cpu().numpy().ravel()
    ev.append(np.where(st != 0, e, 0)); nv.append(np.where(st != 0, r["n_valid"].cpu().numpy().ravel(), 0))
ev, nv = np.array(ev), np.array(nv)
m = ev > 0
print("evaluations per refinement: mean %.2f, per-step max mean %.2f (min %d max %d); histogram %s" % (ev[m].mean(), ev.max(axis=1).mean(), ev.max(axis=1).min(), ev.max(axis=1).max(), np.bincount(ev[m])))
print("valid samples: mean %.0f, per-step max mean %.0f" % (nv[m].mean(), nv.max(axis=1).mean()))
w = (ev * nv)
print("samples x evaluations: mean %.0f, per-step max mean %.0f (ratio %.2f)" % (w[m].mean(), w.max(axis=1).mean(), w.max(axis=1).mean() / w[m].mean()))
