"""Synthetic world shared by tests, bench and the oracle checks.

The reference ships no data (april_group.json, CameraParams.npz and the test
videos are git-ignored, reference .gitignore:131-142), so every input on this
path is synthetic: a regular dodecahedron carrying twelve tag36h11 tags
(ids 0..11, JSON key order 0..11), a pinhole camera, smooth Lissajous
trajectories, noisy corner detections and a small ray-cast renderer.

Nothing here is part of the accelerated path.  It only produces inputs: the
JSON the reference's ``PoseDetector.get_extrinsics`` reads
(detect_pose.py:105-145), the camera matrix ``Calibration`` would load
(calibrate_camera.py:107-123) and gray frames.  The numpy renderer below is
the executable specification of the CUDA renderer in ``csrc/agt_render.cu``
(same geometry, same sub-sample pattern, same integer noise hash); it is used
to create the committed golden fixtures in a container without a GPU.
"""
from __future__ import annotations

import json
import math
from dataclasses import dataclass
from pathlib import Path
from typing import Dict, List, Tuple

import numpy as np

# ----------------------------------------------------------------------------
# constants of the synthetic world
# ----------------------------------------------------------------------------
NUM_TAGS = 12
INRADIUS = 0.030            # metres, centre -> face
TAG_SIZE = 0.020            # metres, outer edge of the black border
CELLS = 10                  # 6x6 data + 1 black border + 1 white quiet zone, each side
CELL = TAG_SIZE / 8.0       # one tag cell in metres
BLACK, WHITE, BACKGROUND = 20.0, 235.0, 128.0
NOISE_SIGMA = 2.0
# tag36h11 ids 0..11, 36 data bits, row-major from the top-left cell, 1 = white.
# Dumped from cv2.aruco.getPredefinedDictionary(DICT_APRILTAG_36h11).
TAG36H11_CODES = (
    0x21A146BAB, 0x92D18FE9B, 0x7089014BB, 0x193979E27, 0x44153D3D7, 0x35CD5B8CF,
    0xA10BA56A0, 0x2B874A608, 0xB57FB8D44, 0x4E20B5A64, 0x61D897F2C, 0xAB3469FFC,
)
# renderer sub-sample pattern: 4x4 taps at +-0.25/+-0.75 px, Gaussian sigma 0.5 px
SUB_OFFSETS = (-0.75, -0.25, 0.25, 0.75)
SUB_SIGMA = 0.5
# dense-refinement surface model
MODEL_GRID = 41
MODEL_HALF_EXTENT = 0.6     # in units of the tag size
MODEL_BLUR_CELLS = 0.2      # Gaussian sigma of the target texture, in cells
MODEL_RASTER = 8            # raster pixels per cell used to pre-blur the texture


@dataclass(frozen=True)
class Camera:
    width: int
    height: int
    fx: float
    fy: float
    cx: float
    cy: float

    @property
    def mtx(self) -> np.ndarray:
        return np.array([[self.fx, 0.0, self.cx], [0.0, self.fy, self.cy], [0.0, 0.0, 1.0]])


CAMERA_VGA = Camera(640, 480, 600.0, 600.0, 320.0, 240.0)
CAMERA_1080P = Camera(1920, 1080, 1400.0, 1400.0, 960.0, 540.0)


# ----------------------------------------------------------------------------
# small rotation helpers (float64; cv2-free)
# ----------------------------------------------------------------------------
def rodrigues(rvec) -> np.ndarray:
    """Axis-angle (3,) -> rotation matrix (3,3), float64."""
    r = np.asarray(rvec, dtype=np.float64).reshape(3)
    theta = math.sqrt(float(r @ r))
    if theta < 1e-300:
        return np.eye(3)
    k = r / theta
    c, s = math.cos(theta), math.sin(theta)
    kx = np.array([[0.0, -k[2], k[1]], [k[2], 0.0, -k[0]], [-k[1], k[0], 0.0]])
    return c * np.eye(3) + (1.0 - c) * np.outer(k, k) + s * kx


def rotation_to_rvec(rmat) -> np.ndarray:
    """Rotation matrix (3,3) -> axis-angle (3,), float64, angle in [0, pi]."""
    m = np.asarray(rmat, dtype=np.float64)
    w = np.array([m[2, 1] - m[1, 2], m[0, 2] - m[2, 0], m[1, 0] - m[0, 1]])
    s = 0.5 * math.sqrt(float(w @ w))
    c = max(-1.0, min(1.0, 0.5 * (np.trace(m) - 1.0)))
    theta = math.atan2(s, c)
    if s > 1e-9:
        return w * (0.5 * theta / s)
    if c > 0.0:
        return 0.5 * w
    # angle ~ pi: recover the axis from the symmetric part
    d = np.sqrt(np.maximum((np.diag(m) + 1.0) * 0.5, 0.0))
    if m[0, 1] < 0:
        d[1] = -d[1]
    if m[0, 2] < 0:
        d[2] = -d[2]
    return theta * d / max(np.linalg.norm(d), 1e-300)


# ----------------------------------------------------------------------------
# AprilGroup geometry
# ----------------------------------------------------------------------------
def face_normals() -> np.ndarray:
    """12 unit face normals: cyclic permutations of (0, +-1, +-phi)."""
    phi = (1.0 + math.sqrt(5.0)) / 2.0
    out = []
    for perm in range(3):
        for s1 in (1.0, -1.0):
            for s2 in (1.0, -1.0):
                v = np.roll(np.array([0.0, s1, s2 * phi]), perm)
                out.append(v / np.linalg.norm(v))
    return np.array(out)


def tag_extrinsics() -> List[Tuple[float, np.ndarray, np.ndarray]]:
    """Per tag (size, tvec(3,), rvec(3,)); rvec rotates +z onto the face normal."""
    z = np.array([0.0, 0.0, 1.0])
    out = []
    for n in face_normals():
        axis = np.cross(z, n)
        s = np.linalg.norm(axis)
        ang = math.atan2(s, float(n[2]))
        rvec = axis / s * ang
        out.append((TAG_SIZE, INRADIUS * n, rvec))
    return out


def april_group_dict() -> Dict:
    """The JSON layout detect_pose.py:122-130 reads: extrinsics[:3]=t, [-3:]=r."""
    tags = {}
    for k, (size, tvec, rvec) in enumerate(tag_extrinsics()):
        tags[str(k)] = {"size": size, "extrinsics": [float(v) for v in tvec] + [float(v) for v in rvec]}
    return {"tags": tags}


def write_april_group_json(root: Path) -> Path:
    """Write april_group.json below ``root`` at the relative path the reference
    opens (detect_pose.py:54-55,115) and return the file path."""
    d = Path(root) / "aprilgroup_tracking" / "aprilgroup_pose_estimation"
    d.mkdir(parents=True, exist_ok=True)
    p = d / "april_group.json"
    p.write_text(json.dumps(april_group_dict(), indent=1))
    return p


def group_transforms_f32() -> Tuple[np.ndarray, np.ndarray]:
    """(R_k (12,3,3), t_k (12,3)) as the reference sees them: rvec/tvec rounded
    to float32 when read from JSON (detect_pose.py:126-130)."""
    rs, ts = [], []
    for _, tvec, rvec in tag_extrinsics():
        rs.append(rodrigues(np.float32(rvec).astype(np.float64)))
        ts.append(np.float32(tvec).astype(np.float64))
    return np.array(rs), np.array(ts)


def object_points() -> np.ndarray:
    """(48,3) float64 corners, index 4*k+j, corner order of transform_helper.py:56-59."""
    r = TAG_SIZE / 2.0
    base = np.array([[-r, -r, 0.0], [-r, r, 0.0], [r, r, 0.0], [r, -r, 0.0]])
    rk, tk = group_transforms_f32()
    return np.concatenate([base @ rk[k].T + tk[k] for k in range(NUM_TAGS)], axis=0)


def tag_cells(tag_id: int) -> np.ndarray:
    """(10,10) float intensities of one tag incl. border and quiet zone; row 0 is +y."""
    g = np.full((CELLS, CELLS), WHITE)
    g[1:9, 1:9] = BLACK
    code = TAG36H11_CODES[tag_id]
    for i in range(36):
        if (code >> (35 - i)) & 1:
            g[2 + i // 6, 2 + i % 6] = WHITE
    return g


def all_tag_cells() -> np.ndarray:
    return np.stack([tag_cells(k) for k in range(NUM_TAGS)])


def _gauss_blur_sep(img: np.ndarray, sigma: float) -> np.ndarray:
    rad = int(math.ceil(4.0 * sigma))
    x = np.arange(-rad, rad + 1, dtype=np.float64)
    k = np.exp(-0.5 * (x / sigma) ** 2)
    k /= k.sum()
    pad = np.pad(img, rad, mode="edge")
    tmp = sum(k[i] * pad[:, i:i + img.shape[1]] for i in range(2 * rad + 1))
    return sum(k[i] * tmp[i:i + img.shape[0], :] for i in range(2 * rad + 1))


def model_pitch() -> float:
    """Metric spacing of neighbouring surface samples."""
    return 2.0 * MODEL_HALF_EXTENT * TAG_SIZE / (MODEL_GRID - 1)


def surface_model() -> Tuple[np.ndarray, np.ndarray, np.ndarray, np.ndarray]:
    """Dense-refinement surface model.

    Returns (samples (S,4) f32 = x,y,z,O in the group frame, sample_tag (S,) u8,
    tag_normals (12,3) f32, tag_centres (12,3) f32), S = 12*41*41, tag-major.
    O is the tag texture pre-blurred by MODEL_BLUR_CELLS and sampled bilinearly.
    """
    rk, tk = group_transforms_f32()
    lin = np.linspace(-MODEL_HALF_EXTENT * TAG_SIZE, MODEL_HALF_EXTENT * TAG_SIZE, MODEL_GRID)
    gx, gy = np.meshgrid(lin, lin)            # gy varies along rows
    samples, tags = [], []
    half = CELLS * CELL / 2.0
    for k in range(NUM_TAGS):
        raster = np.kron(tag_cells(k), np.ones((MODEL_RASTER, MODEL_RASTER)))
        raster = _gauss_blur_sep(raster, MODEL_BLUR_CELLS * MODEL_RASTER)
        # raster pixel centres: col c at x = -half + (c+0.5)*CELL/R ; row r at y = +half - (r+0.5)*CELL/R
        fc = (gx + half) / (CELL / MODEL_RASTER) - 0.5
        fr = (half - gy) / (CELL / MODEL_RASTER) - 0.5
        n = raster.shape[0]
        fc = np.clip(fc, 0.0, n - 1.000001)
        fr = np.clip(fr, 0.0, n - 1.000001)
        c0, r0 = np.floor(fc).astype(int), np.floor(fr).astype(int)
        a, b = fc - c0, fr - r0
        o = ((1 - a) * (1 - b) * raster[r0, c0] + a * (1 - b) * raster[r0, c0 + 1]
             + (1 - a) * b * raster[r0 + 1, c0] + a * b * raster[r0 + 1, c0 + 1])
        local = np.stack([gx.ravel(), gy.ravel(), np.zeros(gx.size)], axis=1)
        pts = local @ rk[k].T + tk[k]
        samples.append(np.concatenate([pts, o.reshape(-1, 1)], axis=1))
        tags.append(np.full(gx.size, k, dtype=np.uint8))
    normals = rk[:, :, 2].copy()
    return (np.concatenate(samples).astype(np.float32), np.concatenate(tags),
            normals.astype(np.float32), tk.astype(np.float32))


# ----------------------------------------------------------------------------
# poses, trajectories, detections
# ----------------------------------------------------------------------------
def random_pose(rng: np.random.Generator) -> np.ndarray:
    """(6,) rvec,tvec inside the working volume of SURVEY.md 8d."""
    rvec = rng.normal(0.0, 0.6, 3)
    tvec = np.array([rng.uniform(-0.10, 0.10), rng.uniform(-0.05, 0.05), rng.uniform(0.25, 0.60)])
    return np.concatenate([rvec, tvec])


def trajectory(seed: int, n_frames: int) -> np.ndarray:
    """(n_frames,6) smooth Lissajous poses: <=0.03 rad and <=3 mm per frame, no
    exactly-zero velocity component (detect_pose.py:236-237 raises on zeros)."""
    rng = np.random.default_rng(seed)
    t = np.arange(n_frames, dtype=np.float64)
    r0 = rng.normal(0.0, 0.4, 3)
    c0 = np.array([rng.uniform(-0.03, 0.03), rng.uniform(-0.02, 0.02), rng.uniform(0.36, 0.48)])
    out = np.zeros((n_frames, 6))
    for i in range(3):
        w = rng.uniform(0.015, 0.035)
        ph = rng.uniform(0.0, 2 * math.pi)
        out[:, i] = r0[i] + 0.35 * np.sin(w * t + ph)                 # <= 0.0123 rad / frame / axis
    amp = (0.06, 0.025, 0.09)
    for i in range(3):
        w = rng.uniform(0.010, 0.022)
        ph = rng.uniform(0.0, 2 * math.pi)
        out[:, 3 + i] = c0[i] + amp[i] * np.sin(w * t + ph)           # <= 2 mm / frame / axis
    return out


def project(points: np.ndarray, pose: np.ndarray, cam: Camera) -> np.ndarray:
    """Pinhole projection (N,3)->(N,2), float64, no distortion."""
    r = rodrigues(pose[:3])
    pc = points @ r.T + pose[3:6]
    return np.stack([cam.fx * pc[:, 0] / pc[:, 2] + cam.cx, cam.fy * pc[:, 1] / pc[:, 2] + cam.cy], axis=1)


def visible_tags(pose: np.ndarray, cos_limit: float = -0.3) -> np.ndarray:
    """Tag ids whose face normal looks at the camera: (R n_k).(c_k/|c_k|) < cos_limit."""
    r = rodrigues(pose[:3])
    rk, tk = group_transforms_f32()
    n_cam = rk[:, :, 2] @ r.T
    c = tk @ r.T + pose[3:6]
    c = c / np.linalg.norm(c, axis=1, keepdims=True)
    return np.nonzero(np.sum(n_cam * c, axis=1) < cos_limit)[0]


def detections(pose: np.ndarray, cam: Camera, rng: np.random.Generator, noise_px: float = 0.1):
    """Synthetic detector output for one frame: list of (tag_id, corners (4,2) f64)."""
    obj = object_points()
    out = []
    for k in visible_tags(pose):
        uv = project(obj[4 * k:4 * k + 4], pose, cam) + rng.normal(0.0, noise_px, (4, 2))
        out.append((int(k), uv))
    return out


# ----------------------------------------------------------------------------
# renderer (numpy specification of csrc/agt_render.cu)
# ----------------------------------------------------------------------------
def _hash32(x: np.ndarray) -> np.ndarray:
    x = x.astype(np.uint32)
    x ^= x >> np.uint32(16)
    x = (x * np.uint32(0x7FEB352D)).astype(np.uint32)
    x ^= x >> np.uint32(15)
    x = (x * np.uint32(0x846CA68B)).astype(np.uint32)
    x ^= x >> np.uint32(16)
    return x


def pixel_noise(seed: int, width: int, height: int) -> np.ndarray:
    """(H,W) float32 approx N(0, NOISE_SIGMA^2): sum of the 4 bytes of a 32-bit hash."""
    idx = np.arange(width * height, dtype=np.uint32)
    with np.errstate(over="ignore"):
        salt = _hash32(np.array([seed & 0xFFFFFFFF], dtype=np.uint32))[0]
        h = _hash32(idx ^ salt)
    s = (h & 0xFF) + ((h >> 8) & 0xFF) + ((h >> 16) & 0xFF) + ((h >> 24) & 0xFF)
    sigma_sum = math.sqrt(4.0 * (256.0 ** 2 - 1.0) / 12.0)
    return ((s.astype(np.float32) - np.float32(510.0)) * np.float32(NOISE_SIGMA / sigma_sum)).reshape(height, width)


def bounding_box(pose: np.ndarray, cam: Camera, margin: float = 3.0) -> Tuple[int, int, int, int]:
    """Pixel box (x0,y0,x1,y1) (inclusive/exclusive) containing the projected body."""
    circ = INRADIUS * 1.2584086 + 1e-4
    tz = max(float(pose[5]) - circ, 1e-3)
    u = cam.fx * pose[3] / pose[5] + cam.cx
    v = cam.fy * pose[4] / pose[5] + cam.cy
    rad = max(cam.fx, cam.fy) * circ / tz * 1.15 + margin
    x0, x1 = int(math.floor(u - rad)), int(math.ceil(u + rad)) + 1
    y0, y1 = int(math.floor(v - rad)), int(math.ceil(v + rad)) + 1
    return max(x0, 0), max(y0, 0), min(x1, cam.width), min(y1, cam.height)


def render(pose: np.ndarray, cam: Camera, seed: int = 0, noise: bool = True) -> np.ndarray:
    """Ray-cast one gray frame (H,W) u8 of the dodecahedron at ``pose``."""
    img = np.full((cam.height, cam.width), BACKGROUND, dtype=np.float32)
    x0, y0, x1, y1 = bounding_box(pose, cam)
    if x1 > x0 and y1 > y0:
        r = rodrigues(pose[:3])
        t = np.asarray(pose[3:6], dtype=np.float64)
        o_obj = -r.T @ t
        rk, tk = group_transforms_f32()
        normals = rk[:, :, 2]
        cells = all_tag_cells()
        xs, ys = np.meshgrid(np.arange(x0, x1, dtype=np.float64), np.arange(y0, y1, dtype=np.float64))
        acc = np.zeros(xs.shape)
        wsum = 0.0
        num = INRADIUS - normals @ o_obj                                  # (12,)
        for oy in SUB_OFFSETS:
            for ox in SUB_OFFSETS:
                w = math.exp(-(ox * ox + oy * oy) / (2 * SUB_SIGMA ** 2))
                d_cam = np.stack([(xs + ox - cam.cx) / cam.fx, (ys + oy - cam.cy) / cam.fy, np.ones_like(xs)], axis=-1)
                d_obj = d_cam @ r                                          # R^T d
                den = d_obj @ normals.T                                    # (h,w,12)
                with np.errstate(divide="ignore", invalid="ignore"):
                    tt = num / den
                t_in = np.where(den < 0, tt, -np.inf)
                t_out = np.where(den > 0, tt, np.inf)
                face = np.argmax(t_in, axis=-1)
                te = np.max(t_in, axis=-1)
                tx = np.min(t_out, axis=-1)
                hit = (te < tx) & (te > 0)
                p = o_obj + te[..., None] * d_obj
                q = np.einsum("hwi,hwij->hwj", p - tk[face], rk[face])     # R_k^T (P - t_k)
                half = CELLS * CELL / 2.0
                col = np.floor((q[..., 0] + half) / CELL).astype(int)
                row = np.floor((half - q[..., 1]) / CELL).astype(int)
                inside = (col >= 0) & (col < CELLS) & (row >= 0) & (row < CELLS)
                colour = np.where(inside, cells[face, np.clip(row, 0, CELLS - 1), np.clip(col, 0, CELLS - 1)], WHITE)
                acc += w * np.where(hit, colour, BACKGROUND)
                wsum += w
        img[y0:y1, x0:x1] = (acc / wsum).astype(np.float32)
    if noise:
        img = img + pixel_noise(seed, cam.width, cam.height)
    return np.clip(np.floor(img + 0.5), 0, 255).astype(np.uint8)
