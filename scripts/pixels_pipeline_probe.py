"""Timeline of the from-pixels stream step with the detector one frame ahead (run on the GPU box): CUDA-event durations of the pieces
on the side stream (ingest copy, detector) and on the main stream (windows, chain), and of each of them alone."""
import sys, numpy as np, torch
sys.path.insert(0, '.')
import bench
from accurate_aprilgroup_tracking_b200 import synth
from accurate_aprilgroup_tracking_b200.batched import BatchedPoseDetector
from accurate_aprilgroup_tracking_b200.context import AgtContext
CAM = bench.CAM
ctx = AgtContext(0, CAM.mtx, None); ctx.set_synthetic_model()
S, F = 64, 24
trajs = [synth.trajectory(5000 + s, F) for s in range(S)]
bank = ctx.alloc_pyramid(S * F, CAM.width, CAM.height, 1)
for f in range(F):
    ctx.render(bank, np.array([trajs[i][f] for i in range(S)]), np.array([1000 * s + f for s in range(S)]), offset=f * S, batch=S)
frames = bank.frames.reshape(F, S, CAM.height, CAM.width)
bpd = BatchedPoseDetector(ctx, S, CAM.width, CAM.height, synth.object_points())
E = lambda: torch.cuda.Event(enable_timing=True)
hp = torch.cuda.Stream(priority=-1)
side, third = torch.cuda.Stream(), torch.cuda.Stream()
stepped, prepared, copied, landed, built = (torch.cuda.Event() for _ in range(5))
seg = {k: [] for k in ("side copy", "side detect", "main windows", "main chain", "main iteration", "third K1")}
with torch.cuda.stream(hp):
    for rep in range(3):
        bpd.reset()
        main = torch.cuda.current_stream()
        bpd.frames.copy_(frames[0]); stepped.record(main)
        evs = []
        for f in range(F):
            e = {k: E() for k in ("i0", "w1", "c0", "c1", "s0", "s1", "s2", "k0", "k1")}
            e["i0"].record(main)
            if f + 1 < F:
                bpd.next_windows(); e["w1"].record(main); prepared.record(main)
                side.wait_event(stepped); side.wait_event(prepared)
                with torch.cuda.stream(side):
                    e["s0"].record(side); bpd.ingest_next(frames[f + 1], build=False); e["s1"].record(side); copied.record(side)
                    bpd.detect_next(); e["s2"].record(side); landed.record(side)
                third.wait_event(copied)
                with torch.cuda.stream(third):
                    e["k0"].record(third); bpd.build_next(); e["k1"].record(third); built.record(third)
            e["c0"].record(main)
            out = bpd.step_frames() if f == 0 else bpd.step(None)
            e["c1"].record(main)
            stepped.record(main)
            if f + 1 < F:
                main.wait_event(landed); main.wait_event(built)
            evs.append(e)
        torch.cuda.synchronize()
        if rep == 2:
            for f in range(4, F - 1):
                e, n = evs[f], evs[f + 1]
                seg["side copy"].append(e["s0"].elapsed_time(e["s1"])); seg["side detect"].append(e["s1"].elapsed_time(e["s2"]))
                seg["main windows"].append(e["i0"].elapsed_time(e["w1"])); seg["main chain"].append(e["c0"].elapsed_time(e["c1"]))
                seg["main iteration"].append(e["i0"].elapsed_time(n["i0"])); seg["third K1"].append(e["k0"].elapsed_time(e["k1"]))
            # offsets inside an iteration, relative to its start on the main stream
            f = 10; e = evs[f]
            print("iteration 10, ms after its start:", {k: round(e["i0"].elapsed_time(e[k]), 3) for k in ("w1", "s0", "s1", "s2", "k0", "k1", "c0", "c1")})
for k, v in seg.items():
    print(f"{k:16s} {1e3 * np.mean(v):7.1f} us (min {1e3 * np.min(v):.1f}, max {1e3 * np.max(v):.1f})")
