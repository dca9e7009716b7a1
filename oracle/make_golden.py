"""Generate the committed golden vectors under tests/golden/ (run in the build container).

    python -m oracle.make_golden

TEST INFRASTRUCTURE.  The reference holds no golden vectors, KATs or fixtures of its own
(no tests, data files git-ignored), so the vectors are produced by
running the real thing here and are committed with this script:

  ape_sequence.npz   the UNMODIFIED reference (/root/reference, via oracle/ref_runner.py) driven over a
                     seeded 80-frame detection sequence with a tracking loss and a gate failure; per-frame
                     prev_transform / extrinsic_guess snapshots + all_objpts
  lk_pair.npz        cv2.calcOpticalFlowPyrLK + cv2.pyrDown + cv2.Scharr (OpenCV 4.13.0) on a rendered
                     320x240 frame pair
  dpr_case.npz       oracle/dpr_oracle.py on a rendered 320x240 frame (no reference code exists for this stage)
  tag_case.npz       cv2.aruco.ArucoDetector (DICT_APRILTAG_36h11) on three rendered 640x480 frames: ids and corners in the
                     reference's corner order, next to the true corners (row N3; the reference's apriltag library is absent)

Versions are recorded inside each file.
"""
from __future__ import annotations

import sys
from pathlib import Path

import cv2
import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from accurate_aprilgroup_tracking_b200 import synth  # noqa: E402
from oracle import ape_oracle, dpr_oracle, lk_oracle, ref_runner  # noqa: E402

GOLDEN = ROOT / "tests" / "golden"
SMALL_CAM = synth.Camera(320, 240, 420.0, 420.0, 160.0, 120.0)
VERSIONS = np.array([f"cv2={cv2.__version__}", f"numpy={np.__version__}"])


def ape_sequence(n_frames: int = 80, seed: int = 1000):
    """Seeded detections incl. one tracking loss (frames 30-31) and one gate failure (frame 50)."""
    cam = synth.CAMERA_VGA
    traj = synth.trajectory(seed, n_frames)
    rng = np.random.default_rng(seed)
    frames = []
    for f in range(n_frames):
        dets = synth.detections(traj[f], cam, rng)
        if f in (30, 31):
            dets = dets[:1]
        if f == 50:
            dets = [(t, c + (30.0 if i == 0 else 0.0)) for i, (t, c) in enumerate(dets)]
        frames.append(dets)
    return cam, traj, frames


def pack_detections(frames):
    ids = np.full((len(frames), 12), -1, np.int32)
    corners = np.zeros((len(frames), 12, 4, 2), np.float64)
    for f, dets in enumerate(frames):
        for j, (t, c) in enumerate(dets):
            ids[f, j] = t
            corners[f, j] = c
    return ids, corners


def unpack_detections(ids, corners, f):
    return [(int(ids[f, j]), corners[f, j]) for j in range(ids.shape[1]) if ids[f, j] >= 0]


def make_ape():
    if not ref_runner.reference_available():
        raise SystemExit("/root/reference is not mounted: ape_sequence.npz can only be generated in the build container")
    cam, traj, frames = ape_sequence()
    rr = ref_runner.ReferenceRunner(cam.mtx, None, True)
    n = len(frames)
    prev = np.full((n, 6), np.nan)
    guess = np.full((n, 6), np.nan)
    nvel = np.zeros(n, np.int32)
    for f, dets in enumerate(frames):
        snap = rr.estimate(dets)
        if snap["prev"] is not None:
            prev[f] = np.concatenate(snap["prev"])
        if snap["guess"] is not None:
            guess[f] = np.concatenate(snap["guess"])
        nvel[f] = snap["n_vel"]
    ids, corners = pack_detections(frames)
    np.savez_compressed(GOLDEN / "ape_sequence.npz", ids=ids, corners=corners, prev=prev, guess=guess, n_vel=nvel,
                        all_objpts=rr.det.all_objpts, mtx=cam.mtx, truth=traj, versions=VERSIONS,
                        source=np.array(["unmodified reference PoseDetector._estimate_pose (detect_pose.py:467-574)"]))


def make_lk():
    cam = SMALL_CAM
    traj = synth.trajectory(3000, 12)
    traj[:, 3:5] *= 0.5
    a = synth.render(traj[10], cam, seed=10)
    b = synth.render(traj[11], cam, seed=11)
    pts = synth.project(synth.object_points(), traj[10], cam).astype(np.float32)
    extra = np.array([[5, 5], [316.5, 3.2], [-3, 10], [400, 100], [160, 239.5], [0, 0], [319, 239], [-30, -30]], np.float32)
    pts = np.concatenate([pts, extra])
    nxt, st, err = lk_oracle.lk_cv(a, b, pts)
    pyr = lk_oracle.pyramid_cv(a, 4)
    np.savez_compressed(GOLDEN / "lk_pair.npz", prev=a, next=b, pts=pts, next_pts=nxt, status=st, err=err,
                        level1=pyr[1], level2=pyr[2], level3=pyr[3], scharr0=lk_oracle.scharr_cv(a),
                        scharr2=lk_oracle.scharr_cv(pyr[2]), versions=VERSIONS,
                        source=np.array(["cv2.calcOpticalFlowPyrLK / cv2.pyrDown / cv2.Scharr defaults"]))


def make_dpr():
    cam = SMALL_CAM
    rng = np.random.default_rng(4242)
    poses, inits, outs, frames = [], [], [], []
    s, tg, n, c = synth.surface_model()
    model = dpr_oracle.Model(s, tg, n, c, synth.model_pitch())
    for i, z in enumerate((0.115, 0.24)):           # level 1 and level 0 at this focal length
        pose = synth.random_pose(rng)
        pose[3:] = (0.004 * (i + 1), -0.003, z)
        frame = synth.render(pose, cam, seed=100 + i)
        init = pose + np.concatenate([rng.normal(0, 0.008, 3), rng.normal(0, 0.0003, 3)])
        out = dpr_oracle.refine(lk_oracle.pyramid_cv(frame, 4), model, cam.mtx, init)
        poses.append(pose); inits.append(init); frames.append(frame)
        outs.append(np.concatenate([out["pose"], [out["cost"], out["n_valid"], out["evals"], out["status"], out["level"]]]))
    np.savez_compressed(GOLDEN / "dpr_case.npz", frames=np.stack(frames), truth=np.array(poses), init=np.array(inits),
                        result=np.array(outs), mtx=cam.mtx, versions=VERSIONS,
                        source=np.array(["oracle/dpr_oracle.py (executable spec; no reference code exists)"]))


def make_undistort():
    """The reference's frame ingest on a small BGR frame: undistort_frame (detect_pose.py:147-183) + BGR2GRAY (:602),
    computed by the calls the reference makes (cv2), for two lens models."""
    import cv2
    w, h, f = 160, 120, 150.0
    rng = np.random.default_rng(777)
    mtx = np.array([[f, 0, w / 2 + 1.7], [0, f * 1.01, h / 2 - 0.9], [0, 0, 1]])
    frame = cv2.GaussianBlur(rng.integers(0, 256, (h, w, 3), dtype=np.uint8), (0, 0), 1.2)
    dists, news, rois, grays = [], [], [], []
    for dist in ([-0.28, 0.11, 0.0007, -0.0004, -0.02], [0.05, -0.1, 0.001, 0.002, 0.03]):
        dist = np.array(dist, np.float64).reshape(1, 5)
        new_mtx, roi = cv2.getOptimalNewCameraMatrix(mtx, dist, (w, h), 1, (w, h))
        x, y, rw, rh = roi
        gray = cv2.cvtColor(cv2.undistort(frame, mtx, dist, None, new_mtx)[y:y + rh, x:x + rw], cv2.COLOR_BGR2GRAY)
        dists.append(dist); news.append(new_mtx); rois.append(np.array(roi)); grays.append(gray.ravel())
    np.savez_compressed(GOLDEN / "undistort_case.npz", frame=frame, mtx=mtx, dist=np.array(dists), new_mtx=np.array(news),
                        roi=np.array(rois), gray0=grays[0], gray1=grays[1], versions=VERSIONS,
                        source=np.array(["cv2.getOptimalNewCameraMatrix + cv2.undistort + cv2.cvtColor, the calls of detect_pose.py:167-177, 602"]))


def make_tags():
    from oracle import tag_oracle
    cam = synth.CAMERA_VGA
    frames, poses, ids, corners, counts = [], [], [], [], []
    for s in (700, 702, 709):
        pose = synth.trajectory(s, 1)[0]
        frame = synth.render(pose, cam, seed=s)
        det = sorted(tag_oracle.detect_cv(frame), key=lambda d: d[0])
        frames.append(frame); poses.append(pose); counts.append(len(det))
        ids.extend(i for i, _ in det); corners.extend(c for _, c in det)
    np.savez_compressed(GOLDEN / "tag_case.npz", frames=np.stack(frames), poses=np.array(poses), counts=np.array(counts),
                        ids=np.array(ids, np.int32), corners=np.array(corners), mtx=cam.mtx, versions=VERSIONS,
                        source=np.array(["cv2.aruco.ArucoDetector(DICT_APRILTAG_36h11).detectMarkers, default parameters; "
                                         "corners rolled to the order of transform_helper.py:56-59"]))


def main():
    GOLDEN.mkdir(parents=True, exist_ok=True)
    make_ape()
    make_lk()
    make_dpr()
    make_undistort()
    make_tags()
    for p in sorted(GOLDEN.glob("*.npz")):
        print(p.name, p.stat().st_size)


if __name__ == "__main__":
    main()
