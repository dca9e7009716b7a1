import sys, numpy as np
sys.path.insert(0, '.')
import cv2
from accurate_aprilgroup_tracking_b200 import synth
from accurate_aprilgroup_tracking_b200.context import AgtContext
from oracle import lk_oracle, dpr_oracle
from tests import util
np.set_printoptions(precision=5, suppress=True, linewidth=200)
cam = synth.CAMERA_VGA
ctx = AgtContext(0, cam.mtx, None); ctx.set_synthetic_model()
# ---- PnP
from tests.test_gpu_kernels import _pnp_inputs
obj, img, valid, truth = _pnp_inputs(cam, 16, 42)
pose, ok, err, iters = [x.cpu().numpy() for x in ctx.pnp(obj, img, valid)]
print("PNP noguess ok", ok, "iters", iters, "err", err)
for i in range(4):
    m = valid[i]==1
    _, r, t = cv2.solvePnP(obj[m], img[i][m], cam.mtx, None, flags=cv2.SOLVEPNP_ITERATIVE)
    print(i, m.sum(), "gpu", pose[i], "cv", r.ravel(), t.ravel(), "truth", truth[i])
rng = np.random.default_rng(7)
guess = truth + np.concatenate([rng.normal(0, 0.03, (16, 3)), rng.normal(0, 0.003, (16, 3))], axis=1)
pose, ok, err, iters = [x.cpu().numpy() for x in ctx.pnp(obj, img, valid, guess, np.ones(16, np.uint8))]
print("PNP guess ok", ok, "iters", iters, "err", err)
for i in range(4):
    print(i, "gpu", pose[i], "truth", truth[i])
# ---- LK
traj = synth.trajectory(3000, 3)
def rend(poses, seeds):
    pyr = ctx.alloc_pyramid(len(poses), cam.width, cam.height, 4)
    ctx.render(pyr, np.asarray(poses), np.asarray(seeds)); ctx.build_pyramid(pyr); ctx.sync(); return pyr
prev = rend(traj[:2], [1,2]); nxt = rend(traj[1:3], [2,3])
pts = np.stack([synth.project(synth.object_points(), traj[i], cam) for i in range(2)]).astype(np.float32)
out, st, er = [x.cpu().numpy() for x in ctx.lk(prev, nxt, pts)]
ro, rs, re = lk_oracle.lk_cv(prev.frames[0].cpu().numpy(), nxt.frames[0].cpu().numpy(), pts[0])
print("LK gpu st", st[0][:16], "cv st", rs[:16])
print("LK gpu out", out[0][:6].ravel(), "cv", ro[:6].ravel(), "in", pts[0][:6].ravel())
print("LK err", er[0][:6], re[:6])
# ---- DPR
cam2 = synth.CAMERA_1080P
ctx2 = AgtContext(0, cam2.mtx, None); ctx2.set_synthetic_model()
rng = np.random.default_rng(2000)
truth = np.array([synth.random_pose(rng) for _ in range(4)])
pyr = ctx2.alloc_pyramid(4, cam2.width, cam2.height, 4)
ctx2.render(pyr, truth, np.arange(4)+2000); ctx2.build_pyramid(pyr); ctx2.sync()
init = truth[:, None, :] + np.concatenate([rng.normal(0, 0.01, (4, 1, 3)), rng.normal(0, 0.0005, (4, 1, 3))], axis=2)
res = {k: v.cpu().numpy() for k, v in ctx2.refine(pyr, init, 1).items()}
model = util.dpr_model()
for b in range(4):
    lv = [pyr.level(l)[b].cpu().numpy() for l in range(4)]
    ref = dpr_oracle.refine(lv, model, cam2.mtx, init[b,0])
    print("DPR", b, "gpu pose", res["pose"][b,0], "cost", res["cost"][b,0], "n", res["n_valid"][b,0], "ev", res["evals"][b,0], "st", res["status"][b,0])
    print("     ref pose", ref["pose"], "cost", ref["cost"], "n", ref["n_valid"], "ev", ref["evals"], "st", ref["status"], "lvl", ref["level"])
    ref1 = dpr_oracle.refine(lv, model, cam2.mtx, init[b,0], max_evals=1)
    res1 = None
print("init", init[:,0])
