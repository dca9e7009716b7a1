"""Build a variant of libagt.so with extra -D flags into scripts/build/ (A/B measurements: AGT_LIBRARY=<path> python bench.py ...).
    python scripts/build_variant.py libagt_roll.so -DAGT_LK_ROLL_RUNS=1"""
import subprocess, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from accurate_aprilgroup_tracking_b200 import _build
out = ROOT / "scripts" / "build" / sys.argv[1]
out.parent.mkdir(exist_ok=True)
flags = [f for f in _build.NVCC_FLAGS if f not in ("-Xptxas", "-v")]
srcs = [str(_build.CSRC / s) for s in _build.SOURCES]
subprocess.check_call([_build._nvcc(), *flags, *sys.argv[2:], "-I", str(_build.INCLUDE), "-I", str(_build.CSRC), "-shared", "-o", str(out), *srcs, "-lcudart"])
print(out)
