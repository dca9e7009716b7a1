"""Shared helpers for the parity tests."""
import math

import numpy as np
from pathlib import Path

from accurate_aprilgroup_tracking_b200 import synth

ROOT = Path(__file__).resolve().parent.parent
ROT_TOL = 1e-4                       # rad   (BASELINE.json north_star)
TRANS_TOL = 1e-3 * synth.TAG_SIZE    # 1e-3 of the tag size = 20 um
FLOW_TOL = 0.01                      # px


def rot_angle(rvec_a, rvec_b) -> float:
    ra, rb = synth.rodrigues(rvec_a), synth.rodrigues(rvec_b)
    m = ra @ rb.T
    c = max(-1.0, min(1.0, 0.5 * (np.trace(m) - 1.0)))
    s = 0.5 * math.sqrt((m[2, 1] - m[1, 2]) ** 2 + (m[0, 2] - m[2, 0]) ** 2 + (m[1, 0] - m[0, 1]) ** 2)
    return math.atan2(s, c)


def pose_diff(a, b):
    a, b = np.asarray(a, dtype=np.float64).reshape(6), np.asarray(b, dtype=np.float64).reshape(6)
    return rot_angle(a[:3], b[:3]), float(np.linalg.norm(a[3:] - b[3:]))


def assert_pose_close(a, b, what=""):
    dr, dt = pose_diff(a, b)
    assert dr <= ROT_TOL, f"{what}: rotation differs by {dr:.3e} rad (tol {ROT_TOL})"
    assert dt <= TRANS_TOL, f"{what}: translation differs by {dt:.3e} m (tol {TRANS_TOL})"


def dpr_model():
    from oracle import dpr_oracle
    s, tg, n, c = synth.surface_model()
    return dpr_oracle.Model(s, tg, n, c, synth.model_pitch())
