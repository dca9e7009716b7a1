"""Debug probe: per-level / per-iteration float sums of ONE corner from a -DAGT_LK_DEBUG build of libagt.so, next to the same
numbers from oracle/lk_oracle.py:lk_np (which is bit-identical to cv2).  Usage (on the GPU box):
python scripts/lk_debug_probe.py scripts/build/lk_mismatch_pair395.npz 40"""
import ctypes as C, subprocess, sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from accurate_aprilgroup_tracking_b200 import _build, _lib
dbg = ROOT / "scripts" / "build" / "libagt_dbg.so"
if "--build" in sys.argv or not dbg.exists():
    srcs = [str(_build.CSRC / s) for s in _build.SOURCES]
    subprocess.check_call([_build._nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
                           "--expt-relaxed-constexpr", "-DAGT_LK_DEBUG", "-I", str(_build.INCLUDE), "-I", str(_build.CSRC), "-shared", "-o",
                           str(dbg), *srcs, "-lcudart"])
    if "--build" in sys.argv:
        sys.exit(0)
_lib.LIB_PATH = dbg
from accurate_aprilgroup_tracking_b200.context import AgtContext
from accurate_aprilgroup_tracking_b200 import synth
from oracle import lk_oracle as L
g = np.load(sys.argv[1]); corner = int(sys.argv[2])
h, w = g["prev"].shape
ctx = AgtContext(0, synth.CAMERA_1080P.mtx, None)
pa, pb = ctx.alloc_pyramid(1, w, h, 4), ctx.alloc_pyramid(1, w, h, 4)
ctx.upload_frames(pa, g["prev"][None]); ctx.upload_frames(pb, g["next"][None])
ctx.build_pyramid(pa); ctx.build_pyramid(pb)
lib = _lib.load()
lib.agt_lk_debug_set.argtypes = [C.c_int]; lib.agt_lk_debug_get.argtypes = [C.c_void_p, C.c_int]
lib.agt_lk_debug_set(corner)
out, st, err = [t.cpu().numpy() for t in ctx.lk(pa, pb, g["pts"][None])]
buf = np.zeros(8192, np.float32)
n = lib.agt_lk_debug_get(buf.ctypes.data, 8192)
print("gpu", out[0, corner], out[0, corner].view(np.uint32), "cv", g["cv"][corner], g["cv"][corner].view(np.uint32))
v = buf[:n]; i = 0
print("GPU trace:")
while i < n:
    if v[i] == -1:
        print("  level", int(v[i + 1]), "A", [repr(float(x)) for x in v[i + 2:i + 5]], "p", v[i + 5:i + 7]); i += 7
    else:
        print("    it", int(v[i]), "ib", repr(float(v[i + 1])), repr(float(v[i + 2])), "n", repr(float(v[i + 3])), repr(float(v[i + 4]))); i += 5
log = []
om, ot = L._mismatch_sum_f32, L._tensor_sum_f32
L._mismatch_sum_f32 = lambda p: (log.append(("b", float(om(p)))), om(p))[1]
L._tensor_sum_f32 = lambda p: (log.append(("A", float(ot(p)))), ot(p))[1]
no, ns, ne = L.lk_np(g["prev"], g["next"], g["pts"][corner:corner + 1])
print("lk_np trace (unscaled sums):"); print(log); print("lk_np", no)
