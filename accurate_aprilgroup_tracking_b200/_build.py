"""Build csrc/libagt.so in-tree with nvcc for sm_100a (no JIT cache, no fallback arch).

``python -m accurate_aprilgroup_tracking_b200._build`` or ``__graft_entry__.build()``.
Each .cu is compiled to an object file in parallel, then linked into one shared
library next to the sources so it travels with the tree to the GPU box.
"""
from __future__ import annotations

import concurrent.futures
import hashlib
import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
INCLUDE = PKG.parent / "include"
LIB = CSRC / "libagt.so"
SOURCES = ["agt_api.cu", "agt_pyramid.cu", "agt_undistort.cu", "agt_lk.cu", "agt_pnp.cu", "agt_ape.cu", "agt_dpr.cu", "agt_corner.cu", "agt_tags.cu", "agt_overlay.cu", "agt_render.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xptxas", "-v", "--expt-relaxed-constexpr"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found; libagt.so cannot be built (there is no CPU fallback)")


def _stamp() -> str:
    h = hashlib.sha256()
    for f in sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + [INCLUDE / "agt.h"]):
        h.update(f.name.encode())
        h.update(f.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> Path:
    stamp_file = CSRC / "build" / "stamp"
    stamp = _stamp()
    if not force and LIB.exists() and stamp_file.exists() and stamp_file.read_text() == stamp:
        return LIB
    nvcc = _nvcc()
    obj_dir = CSRC / "build"
    obj_dir.mkdir(exist_ok=True)

    def compile_one(src: str):
        obj = obj_dir / (src[:-3] + ".o")
        cmd = [nvcc, *NVCC_FLAGS, "-I", str(INCLUDE), "-I", str(CSRC), "-c", str(CSRC / src), "-o", str(obj)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        return src, obj, r

    objs = []
    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as ex:
        for src, obj, r in ex.map(compile_one, SOURCES):
            (obj_dir / (src[:-3] + ".ptxas.log")).write_text(r.stderr)
            if r.returncode != 0:
                raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
            if verbose:
                print(r.stderr, file=sys.stderr)
            objs.append(str(obj))
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", str(LIB), *objs, "-lcudart"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    stamp_file.write_text(stamp)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
