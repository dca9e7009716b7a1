import sys, numpy as np, torch
sys.path.insert(0, '.')
import bench
from accurate_aprilgroup_tracking_b200.context import AgtContext
ctx = AgtContext(0, bench.CAM.mtx, None); ctx.set_synthetic_model()
B = 4096
pyr = ctx.alloc_pyramid(B, bench.CAM.width, bench.CAM.height, 4)
for rank in range(8):
    seed = 2000 + 7919 * rank
    truth, init = bench.make_poses(B, seed)
    for b0 in range(0, B, 512):
        ctx.render(pyr, truth[b0:b0 + 512], np.arange(b0, b0 + 512) + seed, offset=b0, batch=512)
    d_init = torch.as_tensor(init, dtype=torch.float64, device=ctx.tdev).reshape(B, 1, 6)
    for _ in range(2): res = ctx.refine(pyr, d_init, 1, fused=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): res = ctx.refine(pyr, d_init, 1, fused=True)
    e1.record(); torch.cuda.synchronize()
    work = float((res["evals"].double() * res["n_valid"].double()).sum())
    print(f"rank {rank} seed {seed}: {e0.elapsed_time(e1) / 5:.3f} ms, mean evals {float(res['evals'].float().mean()):.2f}, sample-evals {work:.3e}")
