"""GPU-vs-oracle parity sweep of dense refinement over many 1080p frames (run on the GPU box)."""
import os, sys, time, numpy as np
sys.path.insert(0, '.')
import torch
from accurate_aprilgroup_tracking_b200 import synth
from accurate_aprilgroup_tracking_b200.context import AgtContext
from oracle import dpr_oracle
from tests import util
import multiprocessing as mp
N = int(sys.argv[1]) if len(sys.argv) > 1 else 512
cam = synth.CAMERA_1080P
ctx = AgtContext(0, cam.mtx, None); ctx.set_synthetic_model()
rng = np.random.default_rng(31337)
truth = np.array([synth.random_pose(rng) for _ in range(N)])
init = truth + np.concatenate([rng.normal(0, 0.01, (N, 3)), rng.normal(0, 0.0005, (N, 3))], axis=1)
pyr = ctx.alloc_pyramid(N, cam.width, cam.height, 4)
ctx.render(pyr, truth, np.arange(N) + 31337); ctx.build_pyramid(pyr); ctx.sync()
model = util.dpr_model()
def work(b):
    lv = [pyr_host[l][b] for l in range(4)]
    r = dpr_oracle.refine(lv, model, cam.mtx, init[b])
    return r["pose"], r["evals"], r["cost"], r["n_valid"]
pyr_host = [pyr.level(l).cpu().numpy() for l in range(4)]
t0 = time.time()
with mp.get_context("fork").Pool(min(32, os.cpu_count())) as pool:
    ref = pool.map(work, range(N))
print("oracle time", time.time() - t0, "cpus", os.cpu_count())
for mode in ("current",):
    for _ in range(2):
        res = ctx.refine(pyr, init.reshape(N, 1, 6), 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); res = ctx.refine(pyr, init.reshape(N, 1, 6), 1); e1.record(); torch.cuda.synchronize()
    out = {k: v.cpu().numpy() for k, v in res.items()}
    dr = np.array([util.pose_diff(out["pose"][b, 0], ref[b][0])[0] for b in range(N)])
    dt = np.array([util.pose_diff(out["pose"][b, 0], ref[b][0])[1] for b in range(N)])
    ev = np.array([r[1] for r in ref]); evg = out["evals"][:, 0]
    cr = np.array([r[2] for r in ref])
    print(f"mode {mode}: kernel {e0.elapsed_time(e1):.3f} ms for {N} poses; rot diff max {dr.max():.2e} p99 {np.percentile(dr,99):.2e} p90 {np.percentile(dr,90):.2e} median {np.median(dr):.2e}; "
          f"trans max {dt.max():.2e}; evals mismatch {(ev != evg).mean():.3f}; >1e-4: {(dr > 1e-4).sum()} >5e-5: {(dr > 5e-5).sum()}; nvalid mismatch {(out['n_valid'][:,0] != np.array([r[3] for r in ref])).sum()}")
    same = ev == evg
    print("   cost rel diff where evals equal: max %.2e median %.2e" % (np.abs(out["cost"][same, 0] - cr[same]).max() / cr.mean(), np.median(np.abs(out["cost"][same, 0] - cr[same]) / cr[same])))
