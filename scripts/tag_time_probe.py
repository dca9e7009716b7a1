"""Per-kernel times of the windowed detector on 64 x 1080p frames (run under ncu --metrics gpu__time_duration.sum)."""
import sys, numpy as np, torch
sys.path.insert(0, '.')
from accurate_aprilgroup_tracking_b200 import synth
from accurate_aprilgroup_tracking_b200.context import AgtContext
cam = synth.CAMERA_1080P
ctx = AgtContext(0, cam.mtx, None)
n = 64
poses = np.array([synth.trajectory(5000 + s, 1)[0] for s in range(n)])
pyr = ctx.alloc_pyramid(n, cam.width, cam.height, 1)
ctx.render(pyr, poses, np.arange(n))
state = ctx.new_stream_state(n)
st = state.cpu().numpy(); st[:, 0] = 1.0; st[:, 1:7] = poses
state.copy_(torch.tensor(st, device=state.device))
radius = float(np.linalg.norm(synth.object_points(), axis=1).max())
rects = ctx.track_rects(state, cam.width, cam.height, radius, 48)
modes = (("window", rects, "auto"), ("whole", None, "auto")) if len(sys.argv) < 2 else \
    (("window", rects, "window"), ("window, local white level", rects, "local"), ("whole, one threshold", None, "window"), ("whole, local white level", None, "local"))
for which, r, mode in modes:
    ctx.set_tag_threshold(mode)
    for _ in range(3): det = ctx.detect_tags(pyr, rects=r)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): det = ctx.detect_tags(pyr, rects=r)
    e1.record(); torch.cuda.synchronize()
    print(f"{which}: {e0.elapsed_time(e1) / 10:.3f} ms per {n} frames, tags found {int(det['n'].sum())}")
area = ((rects[:, 2] - rects[:, 0]) * (rects[:, 3] - rects[:, 1])).float().mean().item()
print(f"mean window {area:.0f} px = {area / (cam.width * cam.height):.3f} of the frame")
