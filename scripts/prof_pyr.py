import sys, numpy as np
sys.path.insert(0, '.')
import torch
from accurate_aprilgroup_tracking_b200 import synth
from accurate_aprilgroup_tracking_b200.context import AgtContext
cam = synth.CAMERA_1080P
B = 256
ctx = AgtContext(0, cam.mtx, None)
pa = ctx.alloc_pyramid(B, cam.width, cam.height, 4)
pa.levels[0].random_(0, 255)
for _ in range(3):
    ctx.build_pyramid(pa)
torch.cuda.synchronize()
print("ok")
