"""Cluster-split refinement (small batches) vs the single-CTA path and vs the oracle (run on the GPU box)."""
import sys, numpy as np
sys.path.insert(0, '.')
if len(sys.argv) > 1:
    from pathlib import Path
    from accurate_aprilgroup_tracking_b200 import _lib
    _lib.LIB_PATH = Path(sys.argv[1]).resolve()
import torch
from accurate_aprilgroup_tracking_b200 import synth
from accurate_aprilgroup_tracking_b200.context import AgtContext
from oracle import dpr_oracle
from tests import util
cam = synth.CAMERA_1080P
ctx = AgtContext(0, cam.mtx, None); ctx.set_synthetic_model()
rng = np.random.default_rng(777)
N = 512
truth = np.array([synth.random_pose(rng) for _ in range(N)])
init = truth + np.concatenate([rng.normal(0, 0.01, (N, 3)), rng.normal(0, 0.0005, (N, 3))], axis=1)
pyr = ctx.alloc_pyramid(N, cam.width, cam.height, 4)
ctx.render(pyr, truth, np.arange(N) + 777); ctx.build_pyramid(pyr); ctx.sync()
big = {k: v.cpu().numpy() for k, v in ctx.refine(pyr, init.reshape(N, 1, 6), 1).items()}          # cluster 1
def timeit(fn, n=20):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for nb in (8, 16, 32, 64, 128, 148, 256):
    sub = ctx.alloc_pyramid(nb, cam.width, cam.height, 4)
    for l in range(4): sub.levels[l].copy_(pyr.levels[l][:nb])
    res = {k: v.cpu().numpy() for k, v in ctx.refine(sub, init[:nb].reshape(nb, 1, 6), 1).items()}
    dr = max(util.pose_diff(res["pose"][b, 0], big["pose"][b, 0])[0] for b in range(nb))
    ev = (res["evals"][:, 0] != big["evals"][:nb, 0]).mean()
    ms = timeit(lambda: ctx.refine(sub, init[:nb].reshape(nb, 1, 6), 1))
    print(f"batch {nb:4d}: {ms*1e3:8.1f} us per launch ({ms*1e3/nb:7.2f} us/pose); max rot diff vs single-CTA path {dr:.2e}; evals mismatch {ev:.3f}; left_roi {int(res['left_roi'].sum())}")
model = util.dpr_model()
nb = 24
sub = ctx.alloc_pyramid(nb, cam.width, cam.height, 4)
for l in range(4): sub.levels[l].copy_(pyr.levels[l][:nb])
res = {k: v.cpu().numpy() for k, v in ctx.refine(sub, init[:nb].reshape(nb, 1, 6), 1).items()}
worst = 0
for b in range(nb):
    ref = dpr_oracle.refine([pyr.level(l)[b].cpu().numpy() for l in range(4)], model, cam.mtx, init[b])
    worst = max(worst, util.pose_diff(res["pose"][b, 0], ref["pose"])[0])
    assert res["n_valid"][b, 0] == ref["n_valid"]
print("cluster path vs oracle: max rot diff", worst)
