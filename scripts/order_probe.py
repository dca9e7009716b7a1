"""Does the launch order of the refinements matter?  Times agt_refine_fused on the bench batch in the generated order,
sorted by valid samples (known before the launch) and sorted by samples x evaluations (known only afterwards: the bound)."""
import sys, numpy as np, torch
sys.path.insert(0, '.')
import bench
from accurate_aprilgroup_tracking_b200.context import AgtContext
ctx = AgtContext(0, bench.CAM.mtx, None); ctx.set_synthetic_model()
B = 4096
pyr = ctx.alloc_pyramid(B, bench.CAM.width, bench.CAM.height, 4)
seed = 2000
truth, init = bench.make_poses(B, seed)
ids = np.arange(B) + seed

def run(order, label):
    t, i, s = truth[order], init[order], ids[order]
    for b0 in range(0, B, 512):
        ctx.render(pyr, t[b0:b0 + 512], s[b0:b0 + 512], offset=b0, batch=512)
    d_init = torch.as_tensor(i, dtype=torch.float64, device=ctx.tdev).reshape(B, 1, 6)
    for _ in range(2): res = ctx.refine(pyr, d_init, 1, fused=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): res = ctx.refine(pyr, d_init, 1, fused=True)
    e1.record(); torch.cuda.synchronize()
    nv = res["n_valid"].reshape(-1).cpu().numpy().astype(np.int64); ev = res["evals"].reshape(-1).cpu().numpy().astype(np.int64)
    print(f"{label}: {e0.elapsed_time(e1) / 10:.3f} ms  (sample-evals {float((nv * ev).sum()):.3e}, max per pose {int((nv * ev).max())}, mean {float((nv * ev).mean()):.0f})", flush=True)
    inv = np.empty(B, np.int64); inv[order] = np.arange(B)
    return nv[inv], ev[inv]

ident = np.arange(B)
nv, ev = run(ident, "generated order")
run(np.argsort(-nv, kind="stable"), "sorted by valid samples, descending")
run(np.argsort(-(nv * ev), kind="stable"), "sorted by samples x evaluations, descending (bound)")
run(np.argsort(nv, kind="stable"), "sorted by valid samples, ascending (worst case)")
run(ident, "generated order again")
