import os, sys, numpy as np
sys.path.insert(0, '.')
import torch, ctypes as C
from accurate_aprilgroup_tracking_b200 import synth
from accurate_aprilgroup_tracking_b200.context import AgtContext
cam = synth.CAMERA_1080P
B = 1024
ctx = AgtContext(0, cam.mtx, None)
pa = ctx.alloc_pyramid(B, cam.width, cam.height, 4)
pa.levels[0].random_(0, 255)
ctx._use_current_stream()
def timeit(fn, n=10):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
d = pa.desc
for v in range(5):
    os.environ["AGT_PYR_VARIANT"] = str(v)
    out = []
    for l in range(3):
        def f(l=l):
            ctx._check(ctx.lib.agt_pyr_down(ctx.h, C.c_void_p(d.data[l]), d.width[l], d.height[l], d.pitch[l], d.frame_stride[l],
                                            C.c_void_p(d.data[l+1]), d.pitch[l+1], d.frame_stride[l+1], B))
        ms = timeit(f)
        byt = B * (d.width[l]*d.height[l] + d.width[l+1]*d.height[l+1])
        out.append(f"L{l}: {ms:.3f} ms {byt/ms/1e6:.0f} GB/s")
    ms = timeit(lambda: ctx.build_pyramid(pa))
    print(f"variant {v}: " + " | ".join(out) + f" | all {ms:.3f} ms -> {2754000*B/ms/1e6/6550.1:.3f} of peak", flush=True)
