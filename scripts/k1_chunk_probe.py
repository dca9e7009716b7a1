"""K1 on whole frames, per-level launches over chunks of frames instead of over the whole batch: with a chunk whose levels 1-2 fit in
the L2 cache the next level's launch reads them there and HBM sees level 0 once + one write of every level (run on the GPU box).
    python scripts/k1_chunk_probe.py [frames]"""
import copy, ctypes as C, sys
sys.path.insert(0, '.')
import torch
from accurate_aprilgroup_tracking_b200 import synth
from accurate_aprilgroup_tracking_b200.context import AgtContext
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
cam = synth.CAMERA_1080P
def timeit(fn, n=10):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
ctx = AgtContext(0, cam.mtx, None)
pa = ctx.alloc_pyramid(B, cam.width, cam.height, 4)
g = torch.Generator(device="cuda"); g.manual_seed(1)
for b0 in range(0, B, 256):
    nb = min(256, B - b0)
    x = torch.randint(0, 256, (nb, 1, cam.height, cam.width), device="cuda", generator=g, dtype=torch.uint8).float()
    x = torch.nn.functional.avg_pool2d(x, 3, 1, 1)
    pa.frames[b0:b0 + nb].copy_(x[:, 0].round().to(torch.uint8))
bytes_alg = sum(pa.desc.width[l] * pa.desc.height[l] for l in range(4)) * B
t_ref = timeit(lambda: ctx.build_pyramid(pa))
want = [pa.levels[l].clone() for l in range(1, 4)]
print(f"whole batch: {t_ref:.3f} ms ({bytes_alg / t_ref / 1e6 / 6550.1:.3f} of peak)")
def sub(b0):
    d = type(pa.desc)()
    C.memmove(C.byref(d), C.byref(pa.desc), C.sizeof(d))
    for l in range(4):
        d.data[l] = pa.desc.data[l] + b0 * pa.desc.frame_stride[l]
    return d
for chunk in (8, 16, 24, 32, 48, 64, 96, 128, 256):
    descs = [(sub(b0), min(chunk, B - b0)) for b0 in range(0, B, chunk)]
    def run():
        for d, nb in descs:
            ctx._check(ctx.lib.agt_build_pyramid(ctx.h, C.byref(d), nb))
    for l in range(1, 4): pa.levels[l].zero_()
    t = timeit(run, 5)
    same = all(bool(torch.equal(pa.levels[l], want[l - 1])) for l in range(1, 4))
    print(f"chunks of {chunk:4d}: {t:.3f} ms ({bytes_alg / t / 1e6 / 6550.1:.3f} of peak) identical {same}")
    # two streams alternating chunks (the level chain of one chunk under the level-0 pass of the next)
