"""Is the 64-stream frame-step bound by the GPU chain or by the host loop?  Times the bench's sequence loop with and without the
dense refinement, and the host time of one iteration without waiting for the GPU."""
import sys, time, numpy as np, torch
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
bank = ctx.alloc_pyramid(S * F, CAM.width, CAM.height, 1)
for f in range(F):
    ctx.render(bank, np.array([trajs[i][f] for i in range(S)]), np.array([1000 * s + f for s in range(S)]), offset=f * S, batch=S)
dets = []
for f in range(F):
    a, b, c = pack_detections([synth.detections(trajs[i][f], CAM, rngs[i]) for i in range(S)])
    dets.append((torch.as_tensor(a, device=ctx.tdev), torch.as_tensor(b, device=ctx.tdev), torch.as_tensor(c, device=ctx.tdev)))
frames = bank.frames.reshape(F, S, CAM.height, CAM.width)
for dense in (True, False):
    bpd = BatchedPoseDetector(ctx, S, CAM.width, CAM.height, synth.object_points(), use_dense_refine=dense)
    def seq(copy=True):
        bpd.reset()
        for f in range(F):
            if copy: bpd.frames.copy_(frames[f])
            bpd.step(*dets[f])
    for _ in range(2): seq()
    torch.cuda.synchronize()
    for copy in (True, False):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter(); e0.record()
        for _ in range(5): seq(copy)
        e1.record(); t_host = time.perf_counter() - t0
        torch.cuda.synchronize()
        print(f"dense refinement {dense}, ingest copy {copy}: GPU {e0.elapsed_time(e1) / 5 / F * 1e3:7.1f} us per frame-step; host loop issued it in {t_host / 5 / F * 1e6:7.1f} us per frame-step", flush=True)
