"""Where does a 64-stream frame-step go?  Replays the captured step graph alone (same inputs), then with the input copies, then with
the ingest of the next frame on a side stream - each timed on the device over 200 steps."""
import sys, numpy as np, torch
sys.path.insert(0, '.')
import bench
from accurate_aprilgroup_tracking_b200 import synth
from accurate_aprilgroup_tracking_b200.batched import BatchedPoseDetector, pack_detections
from accurate_aprilgroup_tracking_b200.context import AgtContext
CAM = bench.CAM
ctx = AgtContext(0, CAM.mtx, None); ctx.set_synthetic_model()
S, F = 64, 8
trajs = [synth.trajectory(5000 + s, F) for s in range(S)]
rngs = [np.random.default_rng(5000 + s) for s in range(S)]
bank = ctx.alloc_pyramid(S * F, CAM.width, CAM.height, 1)
for f in range(F):
    ctx.render(bank, np.array([trajs[i][f] for i in range(S)]), np.array([1000 * s + f for s in range(S)]), offset=f * S, batch=S)
dets = []
for f in range(F):
    rows = []
    for i in range(S):
        d = synth.detections(trajs[i][f], CAM, rngs[i])
        if (f + 3 * i) % 17 == 16: d = d[:1]
        rows.append(d)
    dets.append(tuple(torch.as_tensor(a, device=ctx.tdev) for a in pack_detections(rows)))
frames = bank.frames.reshape(F, S, CAM.height, CAM.width)
bpd = BatchedPoseDetector(ctx, S, CAM.width, CAM.height, synth.object_points())
for f in range(F):
    bpd.frames.copy_(frames[f]); bpd.step(*dets[f])
torch.cuda.synchronize()
assert all(g is not None for g in bpd._graphs)
N = 200
def timed(fn, label):
    for _ in range(10): fn(0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(N): fn(k)
    e1.record(); torch.cuda.synchronize()
    print(f"{label:70s} {e0.elapsed_time(e1) / N * 1e3:7.1f} us per step", flush=True)
timed(lambda k: bpd._graphs[k % 3].replay(), "graph replay alone (lk, front, prep, dpr, commit)")
def with_copies(k):
    a, b, c = dets[k % F]
    bpd.in_img.copy_(a); bpd.in_valid.copy_(b); bpd.in_ntags.copy_(c)
    bpd._graphs[k % 3].replay()
timed(with_copies, "+ three input copies")
hist = torch.zeros((S, N, 6), dtype=torch.float64, device=ctx.tdev)
def with_hist(k):
    with_copies(k); hist[:, k].copy_(bpd._outs[k % 3]["pose"])
timed(with_hist, "+ pose copy into the history")
side, landed, stepped = torch.cuda.Stream(), torch.cuda.Event(), torch.cuda.Event()
def with_ingest(k):
    main = torch.cuda.current_stream()
    stepped.record(main)
    side.wait_event(stepped)
    with torch.cuda.stream(side):
        bpd.pyr[(k + 1) % 3].frames.copy_(frames[(k + 1) % F]); ctx.build_pyramid(bpd.pyr[(k + 1) % 3]); landed.record(side)
    with_hist(k)
    main.wait_event(landed)
timed(with_ingest, "+ ingest copy and K1 of the next frame on a side stream")
# the pieces of the chain, eagerly, one at a time (events around each call; launch latency included)
prv, cur = bpd.pyr[0], bpd.pyr[1]
def piece(label, fn):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50): fn()
    e1.record(); torch.cuda.synchronize()
    print(f"  {label:40s} {e0.elapsed_time(e1) / 50 * 1e3:7.1f} us back to back", flush=True)
a, b, c = dets[3]
bpd.in_img.copy_(a); bpd.in_valid.copy_(b); bpd.in_ntags.copy_(c)
piece("lk (fallback frames only)", lambda: ctx.lk(prv, cur, bpd.prev_pts, n_tags=bpd.in_ntags))
nxt, st, _ = ctx.lk(prv, cur, bpd.prev_pts, n_tags=bpd.in_ntags)
piece("front (merge, prepare, pnp, gate)", lambda: ctx.streams_front(bpd.obj, bpd.state, True, bpd.in_img, bpd.in_valid, bpd.in_ntags, tracked=nxt, lk_status=st, prev_valid=bpd.prev_valid))
fr = ctx.streams_front(bpd.obj, bpd.state, True, bpd.in_img, bpd.in_valid, bpd.in_ntags, tracked=nxt, lk_status=st, prev_valid=bpd.prev_valid)
piece("refine (prep + dpr<4>)", lambda: ctx.refine(cur, fr["pose"].reshape(S, 1, 6), 1, mask=fr["gate"]))
piece("K1 (3 levels, 64 frames)", lambda: ctx.build_pyramid(cur))
