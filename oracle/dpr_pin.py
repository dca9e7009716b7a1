"""Independent pins of the dense-refinement oracle (oracle/dpr_oracle.py).

TEST INFRASTRUCTURE (see oracle/__init__.py).  The reference has no dense-refinement code, so the
oracle cannot be pinned to it; it is pinned to third-party code that shares nothing with it:

  residual   cv2.projectPoints (projection) + scipy.ndimage.map_coordinates(order=1) (bilinear
             sampling) + cv2.Scharr -> `independent_residuals`, compared with Evaluator.residuals
  Jacobian   the interpolated Scharr gradient times a central-difference derivative of
             cv2.projectPoints under the left perturbation R <- exp(w) R -> `independent_jacobian`
  answer     `independent_fixed_point`: scipy.optimize.least_squares(method="lm") on the residual to
             get near, then scipy.optimize.root(method="hybr", MINPACK hybrd) on g(p) = J^T r.  The
             oracle's loop is Gauss-Newton with the Scharr gradient in place of the interpolant's
             derivative: the pose it contracts onto is the root of g, whatever the solver.

`python -m oracle.dpr_pin` writes tests/golden/dpr_pin.npz: 64 VGA + 64 1080p noisy renders (frames are
regenerated from the stored seeds by synth.render), the pose scipy found and the oracle's answer.
"""
from __future__ import annotations

import sys
from pathlib import Path

import cv2
import numpy as np

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

from accurate_aprilgroup_tracking_b200 import synth  # noqa: E402
from oracle import dpr_oracle, lk_oracle  # noqa: E402

GOLDEN = ROOT / "tests" / "golden" / "dpr_pin.npz"
CAMERAS = {"vga": synth.CAMERA_VGA, "1080p": synth.CAMERA_1080P}
N_PER_CAMERA = 64
PERTURB_ROT, PERTURB_TRANS = 0.01, 0.0005          # BASELINE config 2: N(0, 0.01 rad), N(0, 0.5 mm)


def model():
    s, tg, n, c = synth.surface_model()
    return dpr_oracle.Model(s, tg, n, c, synth.model_pitch())


def case(cam_name: str, i: int):
    """Seeded (truth, init, frame) of pin case i."""
    cam = CAMERAS[cam_name]
    rng = np.random.default_rng(52000 + (0 if cam_name == "vga" else 1000) + i)
    truth = synth.random_pose(rng)
    init = truth + np.concatenate([rng.normal(0, PERTURB_ROT, 3), rng.normal(0, PERTURB_TRANS, 3)])
    frame = synth.render(truth, cam, seed=52000 + i, noise=True)
    return truth, init, frame


def _pose_of(p, r_init, t_init):
    return np.concatenate([dpr_oracle.log_rotation(dpr_oracle.rodrigues(p[:3]) @ r_init), t_init + p[3:]])


def independent_residuals(pyramid, mdl, kmat, pose0, pose):
    """r (n_selected,), valid mask: the spec's residual from OpenCV's projection and scipy's bilinear sampling."""
    from scipy import ndimage
    ev = dpr_oracle.Evaluator(pyramid, mdl, kmat, pose0)             # only for the frozen sample set and level
    lvl = pyramid[ev.level]
    rvec = cv2.Rodrigues(dpr_oracle.rodrigues(pose[:3]))[0]
    uv = cv2.projectPoints(ev.x.reshape(-1, 1, 3), rvec, np.asarray(pose[3:6], float).reshape(3, 1), kmat, None)[0].reshape(-1, 2)
    ul, vl = uv[:, 0] / (1 << ev.level), uv[:, 1] / (1 << ev.level)
    h, w = lvl.shape
    ok = (np.floor(ul) >= 1) & (np.floor(ul) <= w - 3) & (np.floor(vl) >= 1) & (np.floor(vl) <= h - 3)
    val = ndimage.map_coordinates(lvl.astype(np.float64), [vl, ul], order=1, mode="nearest")
    return np.where(ok, val - ev.o, 0.0), ok


def independent_jacobian(pyramid, mdl, kmat, pose0, pose, eps=1e-7):
    """(n,6) [dI/du dI/dv](Scharr, bilinear, per full-resolution pixel) @ d(u,v)/d(w,t) by central differences of cv2.projectPoints."""
    from scipy import ndimage
    ev = dpr_oracle.Evaluator(pyramid, mdl, kmat, pose0)
    lvl = pyramid[ev.level]
    rmat, t = dpr_oracle.rodrigues(pose[:3]), np.asarray(pose[3:6], float)

    def proj(rm, tt):
        return cv2.projectPoints(ev.x.reshape(-1, 1, 3), cv2.Rodrigues(rm)[0], tt.reshape(3, 1), kmat, None)[0].reshape(-1, 2)

    uv = proj(rmat, t)
    ul, vl = uv[:, 0] / (1 << ev.level), uv[:, 1] / (1 << ev.level)
    g = [ndimage.map_coordinates(cv2.Scharr(lvl, cv2.CV_16S, dx, dy).astype(np.float64), [vl, ul], order=1, mode="nearest")
         / 32.0 / (1 << ev.level) for dx, dy in ((1, 0), (0, 1))]
    jac = np.zeros((len(uv), 6))
    for k in range(6):
        d = np.zeros(6)
        d[k] = eps
        up = proj(dpr_oracle.rodrigues(d[:3]) @ rmat, t + d[3:])
        um = proj(dpr_oracle.rodrigues(-d[:3]) @ rmat, t - d[3:])
        duv = (up - um) / (2 * eps)
        jac[:, k] = g[0] * duv[:, 0] + g[1] * duv[:, 1]
    return jac


def independent_fixed_point(pyramid, mdl, kmat, init):
    """Pose with J^T r = 0 found by scipy alone (MINPACK lmder to get near, then MINPACK hybrd on g) -> (pose, |g|_inf, info)."""
    from scipy.optimize import least_squares, root
    ev = dpr_oracle.Evaluator(pyramid, mdl, kmat, init)
    r_init, t_init = dpr_oracle.rodrigues(init[:3]), np.asarray(init[3:6], float)

    def fun(p):
        return ev.residuals(dpr_oracle.rodrigues(p[:3]) @ r_init, t_init + p[3:], want_jac=False)[0]

    def jac(p):
        return ev.residuals(dpr_oracle.rodrigues(p[:3]) @ r_init, t_init + p[3:])[2]

    def g(p):
        r, _, j = ev.residuals(dpr_oracle.rodrigues(p[:3]) @ r_init, t_init + p[3:])
        return j.T @ r

    near = least_squares(fun, np.zeros(6), jac=jac, method="lm", xtol=1e-10, ftol=1e-10, gtol=1e-10, max_nfev=200)
    scale = np.array([1e-3] * 3 + [1e-4] * 3)          # hybrd's forward differences want unknowns of order one
    sol = root(lambda q: g(near.x + q * scale) * scale, np.zeros(6), method="hybr", options={"xtol": 1e-13, "maxfev": 4000})
    p = near.x + sol.x * scale
    return _pose_of(p, r_init, t_init), float(np.abs(g(p)).max()), {"lm_nfev": near.nfev, "root_nfev": sol.nfev, "root_ok": bool(sol.success)}


def smooth_model(blur_cells: float):
    """The surface model with its target texture blurred by `blur_cells` (synth.surface_model's own construction)."""
    old = synth.MODEL_BLUR_CELLS
    try:
        synth.MODEL_BLUR_CELLS = blur_cells
        s, tg, n, c = synth.surface_model()
    finally:
        synth.MODEL_BLUR_CELLS = old
    return dpr_oracle.Model(s, tg, n, c, synth.model_pitch())


def smooth_render(pose, cam, blur_cells: float) -> np.ndarray:
    """Noise-free frame that is consistent with `smooth_model(blur_cells)`: every pixel's ray is cast onto the nearest face and
    the face's blurred texture raster - the very raster the model's target intensities are sampled from - is sampled
    bilinearly at the hit point.  At the true pose the photometric residual is then interpolation error only."""
    import math
    img = np.full((cam.height, cam.width), synth.BACKGROUND, dtype=np.float64)
    x0, y0, x1, y1 = synth.bounding_box(pose, cam)
    r = synth.rodrigues(pose[:3])
    t = np.asarray(pose[3:6], dtype=np.float64)
    o_obj = -r.T @ t
    rk, tk = synth.group_transforms_f32()
    normals = rk[:, :, 2]
    px = synth.MODEL_RASTER
    rasters = np.stack([synth._gauss_blur_sep(np.kron(synth.tag_cells(k), np.ones((px, px))), blur_cells * px) for k in range(synth.NUM_TAGS)])
    xs, ys = np.meshgrid(np.arange(x0, x1, dtype=np.float64), np.arange(y0, y1, dtype=np.float64))
    d_obj = np.stack([(xs - cam.cx) / cam.fx, (ys - cam.cy) / cam.fy, np.ones_like(xs)], axis=-1) @ r
    den = d_obj @ normals.T
    num = synth.INRADIUS - normals @ o_obj
    with np.errstate(divide="ignore", invalid="ignore"):
        tt = num / den
    t_in = np.where(den < 0, tt, -np.inf)
    t_out = np.where(den > 0, tt, np.inf)
    face, te, tx = np.argmax(t_in, axis=-1), np.max(t_in, axis=-1), np.min(t_out, axis=-1)
    hit = (te < tx) & (te > 0)
    q = np.einsum("hwi,hwij->hwj", o_obj + te[..., None] * d_obj - tk[face], rk[face])
    half = synth.CELLS * synth.CELL / 2.0
    n = rasters.shape[1]
    fc = np.clip((q[..., 0] + half) / (synth.CELL / px) - 0.5, 0.0, n - 1.000001)
    fr = np.clip((half - q[..., 1]) / (synth.CELL / px) - 0.5, 0.0, n - 1.000001)
    c0, r0 = np.floor(fc).astype(int), np.floor(fr).astype(int)
    a, b = fc - c0, fr - r0
    val = ((1 - a) * (1 - b) * rasters[face, r0, c0] + a * (1 - b) * rasters[face, r0, c0 + 1]
           + (1 - a) * b * rasters[face, r0 + 1, c0] + a * b * rasters[face, r0 + 1, c0 + 1])
    img[y0:y1, x0:x1] = np.where(hit, val, synth.BACKGROUND)
    return np.clip(np.floor(img + 0.5), 0, 255).astype(np.uint8)


def main():
    mdl = model()
    out = {k: [] for k in ("cam", "index", "truth", "init", "scipy_pose", "scipy_g", "oracle_pose", "oracle_evals", "oracle_status", "level")}
    for cam_name, cam in CAMERAS.items():
        for i in range(N_PER_CAMERA):
            truth, init, frame = case(cam_name, i)
            pyr = lk_oracle.pyramid_cv(frame, 4)
            pose, gmax, info = independent_fixed_point(pyr, mdl, cam.mtx, init)
            ours = dpr_oracle.refine(pyr, mdl, cam.mtx, init)
            m = dpr_oracle.rodrigues(ours["pose"][:3]) @ dpr_oracle.rodrigues(pose[:3]).T
            dr = 0.5 * np.sqrt((m[2, 1] - m[1, 2]) ** 2 + (m[0, 2] - m[2, 0]) ** 2 + (m[1, 0] - m[0, 1]) ** 2)
            dt = np.linalg.norm(ours["pose"][3:] - pose[3:])
            print(f"{cam_name} {i}: oracle vs scipy {dr:.2e} rad {dt:.2e} m; |g| {gmax:.1e} {info}; evals {ours['evals']} status {ours['status']} level {ours['level']}", flush=True)
            for k, v in zip(out, (0 if cam_name == "vga" else 1, i, truth, init, pose, gmax, ours["pose"], ours["evals"], ours["status"], ours["level"])):
                out[k].append(v)
    import scipy
    np.savez_compressed(GOLDEN, **{k: np.array(v) for k, v in out.items()},
                        versions=np.array([f"cv2={cv2.__version__}", f"numpy={np.__version__}", f"scipy={scipy.__version__}"]),
                        source=np.array(["oracle/dpr_pin.py: scipy least_squares(lm) + root(hybr) on J^T r; frames = synth.render(truth, cam, 52000 + i)"]))
    print("wrote", GOLDEN)


if __name__ == "__main__":
    main()
