"""GPU parity tests: every kernel of the path against the CPU oracle, through the C ABI.

Oracles (see oracle/__init__.py): OpenCV 4.13 for pyrDown / Scharr / calcOpticalFlowPyrLK /
solvePnP (the library the reference calls), oracle/ape_oracle.py for the reference's APE
state machine, oracle/dpr_oracle.py for dense refinement.  Frames are rendered on the GPU
and copied back, so both sides see identical inputs.
"""
import numpy as np
import pytest

from accurate_aprilgroup_tracking_b200 import synth
from tests import util

pytestmark = pytest.mark.gpu


def _render(ctx, cam, poses, seeds, levels=4, noise=True):
    pyr = ctx.alloc_pyramid(len(poses), cam.width, cam.height, levels)
    ctx.render(pyr, np.asarray(poses), np.asarray(seeds), noise=noise)
    ctx.build_pyramid(pyr)
    ctx.sync()
    return pyr


# ------------------------------------------------------------------------------------------
# synthetic renderer sanity (not part of the path; only checks the generator against its spec)
# ------------------------------------------------------------------------------------------
def test_render_matches_numpy_spec(ctxvga):
    rng = np.random.default_rng(11)
    pose = synth.random_pose(rng)
    pose[5] = 0.33
    pyr = _render(ctxvga, synth.CAMERA_VGA, [pose], [5])
    gpu = pyr.frames[0].cpu().numpy()
    ref = synth.render(pose, synth.CAMERA_VGA, seed=5)
    d = np.abs(gpu.astype(int) - ref.astype(int))
    assert d.mean() < 0.05 and (d > 2).mean() < 1e-3, (d.mean(), d.max())


# ------------------------------------------------------------------------------------------
# K1 pyramid + Scharr: bit exact vs OpenCV
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("w,h", [(640, 480), (1920, 1080), (637, 479), (333, 201), (64, 48), (17, 9)])
def test_pyramid_bit_exact(ctxvga, w, h):
    import cv2
    torch = ctxvga.torch
    rng = np.random.default_rng(w * 1000 + h)
    frames = rng.integers(0, 256, (3, h, w), dtype=np.uint8)
    frames[1] = cv2.GaussianBlur(frames[1], (0, 0), 2.0)
    pyr = ctxvga.alloc_pyramid(3, w, h, 4)
    ctxvga.upload_frames(pyr, frames)
    ctxvga.build_pyramid(pyr)
    for b in range(3):
        ref = frames[b]
        for l in range(1, 4):
            ref = cv2.pyrDown(ref)
            got = pyr.level(l)[b].cpu().numpy()
            assert got.shape == ref.shape
            assert np.array_equal(got, ref), f"level {l} frame {b}: {np.abs(got.astype(int) - ref.astype(int)).max()}"
    for l in range(4):
        sch = ctxvga.scharr(pyr, l).cpu().numpy()
        for b in range(3):
            lvl = pyr.level(l)[b].cpu().numpy()
            assert np.array_equal(sch[b, :, :, 0], cv2.Scharr(lvl, cv2.CV_16S, 1, 0))
            assert np.array_equal(sch[b, :, :, 1], cv2.Scharr(lvl, cv2.CV_16S, 0, 1))


def test_masked_pyramid_large_batch(ctxvga):
    # masked builds of a large batch run on a small grid whose warps stride over the flagged frames' strips
    # (and leave at once when nothing is flagged): flagged frames equal the full pyramid, the others are untouched
    import cv2
    torch = ctxvga.torch
    n, w, h = 400, 640, 480
    pyr = ctxvga.alloc_pyramid(n, w, h, 4)
    pyr.levels[0].random_(0, 256)
    for l in (1, 2, 3):
        pyr.levels[l].fill_(255)
    mask = torch.zeros(n, dtype=torch.uint8, device=pyr.levels[0].device)
    ctxvga.build_pyramid_masked(pyr, mask)
    assert all(bool((pyr.levels[l] == 255).all()) for l in (1, 2, 3))
    flagged = [0, 7, 131, 399]
    mask[flagged] = 1
    ctxvga.build_pyramid_masked(pyr, mask)
    for b in flagged:
        ref = pyr.frames[b].cpu().numpy()
        for l in (1, 2, 3):
            ref = cv2.pyrDown(ref)
            assert np.array_equal(pyr.level(l)[b].cpu().numpy(), ref), (b, l)
    for b in (1, 6, 8, 130, 398):
        assert all(bool((pyr.level(l)[b] == 255).all()) for l in (1, 2, 3)), b


def test_pyramid_idempotent_constant(ctxvga):
    # size-independent property: a constant image stays constant at every level
    pyr = ctxvga.alloc_pyramid(2, 1920, 1080, 4)
    pyr.levels[0].fill_(77)
    ctxvga.build_pyramid(pyr)
    for l in range(1, 4):
        assert int(pyr.level(l).min()) == 77 and int(pyr.level(l).max()) == 77


# ------------------------------------------------------------------------------------------
# N4: batched pose overlay vs the reference's cv.circle loop
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("w,h", [(1280, 720), (1920, 1080), (333, 201)])
def test_overlay_points_equal_the_cv_circle_loop(ctx1080, w, h):
    """agt_draw_points against draw.py:144-151 run on the host: np.round, the 1280 x 720 bounds test on the centre, cv.circle(img,
    (x, y), 5, (0, 0, 255), -1) - bit-identical frames, including discs cut by the frame border, centres outside the bounds,
    NaN projections and masked (not accepted) frames."""
    import cv2
    import torch
    rng = np.random.default_rng(w)
    n = 5
    cam = synth.Camera(w, h, 1400.0 * w / 1920, 1400.0 * w / 1920, w / 2.0, h / 2.0)
    ctx = ctx1080
    poses = synth.trajectory(5100, n)
    poses[1, 3] += 0.9 * poses[1, 5] * (w / 2.0) / cam.fx          # object on the right border: discs cut by it, centres outside
    poses[2, 4] -= 0.9 * poses[2, 5] * (h / 2.0) / cam.fy
    old = (ctx.mtx.copy(), ctx.dist)
    ctx.set_camera(cam.mtx, None)
    try:
        host = rng.integers(0, 256, (n, h, w, 3), dtype=np.uint8)
        dev = torch.from_numpy(host).to(ctx.tdev)
        mask = np.array([1, 1, 1, 0, 1], np.uint8)
        obj = synth.object_points()
        proj = ctx.overlay_points(dev, obj, poses, mask).cpu().numpy()
        want = host.copy()
        for f in range(n):
            if not mask[f]:
                continue
            for x, y in np.round(proj[f]).astype(int):
                if 0 <= y < 720 and 0 <= x < 1280:
                    cv2.circle(want[f], (int(x), int(y)), 5, (0, 0, 255), -1)
        got = dev.cpu().numpy()
        assert np.array_equal(got, want)
        assert (got != host).any() and np.array_equal(got[3], host[3])
        # projections are the ones cv.projectPoints gives (float32 object points on the device)
        ref = cv2.projectPoints(obj, poses[0, :3], poses[0, 3:], cam.mtx, None)[0].reshape(-1, 2)
        assert np.abs(proj[0] - ref).max() < 1e-2
        # NaN poses draw nothing and do not fault
        bad = poses.copy(); bad[0, :] = np.nan
        dev2 = torch.from_numpy(host).to(ctx.tdev)
        ctx.overlay_points(dev2, obj, bad, mask)
        assert np.array_equal(dev2[0].cpu().numpy(), host[0]) and np.array_equal(dev2[1:].cpu().numpy(), got[1:])
    finally:
        ctx.set_camera(*old)


# ------------------------------------------------------------------------------------------
# N3 (first step): corner refinement vs cv2.cornerSubPix
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("cam_name,win", [("vga", 5), ("vga", 3), ("1080p", 5), ("1080p", 7)])
def test_corner_subpix_matches_opencv(ctxvga, ctx1080, cam_name, win):
    """agt_corner_subpix against cv2.cornerSubPix (the corner refinement of OpenCV's ArUco detector, which stands in for the
    reference's apriltag `refine_edges`): predicted corners = true projections + 1 px of noise, plus points on and
    outside the image border; tolerance 2e-3 px (the two accumulate the same float64 sums in a different order; the
    iteration stops at 1e-3 px)."""
    from oracle import corner_oracle
    ctx, cam = (ctxvga, synth.CAMERA_VGA) if cam_name == "vga" else (ctx1080, synth.CAMERA_1080P)
    n = 6
    traj = synth.trajectory(4400, n)
    pyr = _render(ctx, cam, traj, np.arange(n) + 4400, levels=1)
    rng = np.random.default_rng(44)
    obj = synth.object_points()
    pts = np.stack([synth.project(obj, traj[i], cam) for i in range(n)]) + rng.normal(0, 1.0, (n, 48, 2))
    extra = np.array([[2.3, 3.1], [cam.width - 2.8, cam.height - 3.1], [0.5, 240.0], [320.0, 0.2], [cam.width - 0.6, 100.0],
                      [-5.0, 10.0], [100.0, cam.height + 3.0]], np.float32)
    pts = np.concatenate([pts, np.broadcast_to(extra, (n,) + extra.shape)], axis=1).astype(np.float32)
    valid = np.ones(pts.shape[:2], np.uint8)
    valid[:, 7] = 0
    out = ctx.corner_subpix(pyr, pts, valid, win=win).cpu().numpy()
    frames = pyr.frames.cpu().numpy()
    worst = worst_all = 0.0
    moved = n_stable = n_all = 0
    for i in range(n):
        inside = (pts[i, :, 0] >= 0) & (pts[i, :, 0] < cam.width) & (pts[i, :, 1] >= 0) & (pts[i, :, 1] < cam.height) & (valid[i] == 1)
        p = pts[i][inside]
        want = corner_oracle.corner_subpix_cv(frames[i], p, win)
        # cv2.cornerSubPix stops a linearly converging iteration at a step of 1e-3 px, and on some corners the iteration does not
        # contract at all: there cv2's own answer moves by up to 0.5 px when the input moves by 1e-4 px.  Parity is asserted on the
        # corners where cv2 agrees with itself under that perturbation (the large majority), and reported on all of them.
        wobble = np.maximum(np.abs(corner_oracle.corner_subpix_cv(frames[i], np.clip(p + np.float32(1e-4), 0, None), win) - want).max(axis=1),
                            np.abs(corner_oracle.corner_subpix_cv(frames[i], np.clip(p - np.float32(1e-4), 0, None), win) - want).max(axis=1))
        stable = wobble < 5e-4
        d = np.abs(out[i][inside] - want).max(axis=1)
        worst = max(worst, float(d[stable].max()))
        worst_all = max(worst_all, float(d.max()))
        n_stable += int(stable.sum()); n_all += int(stable.size)
        assert np.array_equal(out[i][~inside], pts[i][~inside])          # invalid / outside points are copied through
        moved += int((np.abs(want - p).max(axis=1) > 0.05).sum())
    print(f"corner_subpix {cam_name} win {win}: worst difference to cv2 {worst:.2e} px on the {n_stable} of {n_all} corners where cv2 is "
          f"stable under a 1e-4 px perturbation ({worst_all:.2e} px on all), {moved} corners moved")
    assert worst <= 2e-3 and moved > 100 and n_stable > 0.8 * n_all


def test_corner_subpix_host_entry_point_refines_in_place(lib_built):
    from accurate_aprilgroup_tracking_b200 import cv_compat
    from oracle import corner_oracle
    cam = synth.CAMERA_VGA
    pose = synth.trajectory(4500, 1)[0]
    img = synth.render(pose, cam, 7)
    pts = (synth.project(synth.object_points(), pose, cam) + 0.7).astype(np.float32).reshape(-1, 1, 2)
    want = corner_oracle.corner_subpix_cv(img, pts, 5)
    got = cv_compat.default_context().cornerSubPix(img, pts, (5, 5), (-1, -1), (3, 30, 0.001))
    assert got is pts and np.abs(pts.reshape(-1, 2) - want).max() <= 2e-3
    with pytest.raises(ValueError):
        cv_compat.default_context().cornerSubPix(img, pts, (5, 3))


# ------------------------------------------------------------------------------------------
# K2 LK vs cv2.calcOpticalFlowPyrLK
# ------------------------------------------------------------------------------------------
def _bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def _negative_weight_point(x, y):
    """A float32 point within the pixel of (x, y) whose level-0 bilinear weights (lk_oracle._weights) have w11 == -1."""
    from oracle import lk_oracle
    f = np.float32
    fx, fy = f(np.floor(x)), f(np.floor(y))
    xs = fx + np.arange(1, 400, dtype=np.float32) * np.spacing(fx)          # every float32 just right of the pixel centre
    ys = fy + np.arange(1, 4000, dtype=np.float32) * np.spacing(fy)
    qx, qy = (xs - f(10)), (ys - f(10))
    a, b = (qx - np.floor(qx))[:, None], (qy - np.floor(qy))[None, :]
    one, sc = f(1), f(16384)
    w00, w01, w10 = np.rint((one - a) * (one - b) * sc), np.rint(a * (one - b) * sc), np.rint((one - a) * b * sc)
    i, j = np.nonzero(16384 - w00 - w01 - w10 < 0)
    assert i.size, "no negative-weight offset in this pixel"
    px, py = xs[i[i.size // 2]], ys[j[i.size // 2]]
    qx, qy = f(px - f(10)), f(py - f(10))
    assert lk_oracle._weights(f(qx - np.floor(qx)), f(qy - np.floor(qy)))[3] < 0
    return px, py


def _lk_case(ctx, cam, seed, n_pairs, extra_pts=None):
    """agt_lk against cv2.calcOpticalFlowPyrLK: status identical, and - because the kernel adds the structure tensor and the
    mismatch vector in OpenCV's float32 order - points and error of every tracked corner BIT-identical (bar: 0.01 px)."""
    from oracle import lk_oracle
    traj = synth.trajectory(seed, n_pairs + 1)
    prev = _render(ctx, cam, traj[:-1], np.arange(n_pairs) + seed)
    nxt = _render(ctx, cam, traj[1:], np.arange(n_pairs) + seed + 1)
    obj = synth.object_points()
    pts = np.stack([synth.project(obj, traj[i], cam) for i in range(n_pairs)]).astype(np.float32)
    # the same corners moved to sub-pixel offsets whose rounded bilinear weights are (w00, w01, w10, -1): the fourth weight of
    # cv::calcOpticalFlowPyrLK is 2^14 minus the other three and goes negative there (signed int16 in OpenCV's pmaddwd)
    snapped = np.stack([[_negative_weight_point(x, y) for x, y in frame[:12]] for frame in pts]).astype(np.float32)
    pts = np.concatenate([pts, snapped], axis=1).astype(np.float32)
    if extra_pts is not None:
        pts = np.concatenate([pts, np.broadcast_to(extra_pts, (n_pairs,) + extra_pts.shape)], axis=1).astype(np.float32)
    out, st, err = ctx.lk(prev, nxt, pts)
    out, st, err = out.cpu().numpy(), st.cpu().numpy(), err.cpu().numpy()
    worst = 0.0
    for i in range(n_pairs):
        a = prev.frames[i].cpu().numpy()
        b = nxt.frames[i].cpu().numpy()
        ro, rs, re = lk_oracle.lk_cv(a, b, pts[i])
        assert np.array_equal(st[i], rs), f"pair {i}: status differs at {np.nonzero(st[i] != rs)[0]}"
        m = rs == 1
        d = np.abs(out[i][m] - ro[m]).max() if m.any() else 0.0
        worst = max(worst, d)
        assert d <= util.FLOW_TOL, f"pair {i}: flow differs by {d} px"
        assert np.array_equal(_bits(out[i][m]), _bits(ro[m])), f"pair {i}: tracked points are not bit-identical to OpenCV (max diff {d} px)"
        assert np.array_equal(_bits(err[i][m]), _bits(re[m])), f"pair {i}: error differs by {np.abs(err[i][m] - re[m]).max()}"
    return worst


def test_lk_vga_with_border_points(ctxvga):
    extra = np.array([[5, 5], [636.5, 3.2], [-3, 10], [700, 100], [320, 479.5], [100, 100], [0, 0], [639, 479],
                      [-30, -30], [639.9, 240.0]], np.float32)
    worst = _lk_case(ctxvga, synth.CAMERA_VGA, 3000, 6, extra)
    print("lk vga worst", worst)


def test_lk_1080p(ctx1080):
    worst = _lk_case(ctx1080, synth.CAMERA_1080P, 3100, 4)
    print("lk 1080p worst", worst)


def test_lk_config3_sweep_512_pairs_every_corner(ctx1080):
    """BASELINE config 3 at its shape (1080p, 4 levels, 21x21, all 48 corners per pair - the hidden tags' corners track
    whatever covers them and are ill-conditioned) on 512 frame pairs: status identical and every tracked corner within
    0.01 px of cv2.calcOpticalFlowPyrLK - in fact bit-identical (round 1 had 3 of 6144 corners above 0.01 px here)."""
    from oracle import lk_oracle
    cam = synth.CAMERA_1080P
    n, chunk = 512, 128
    obj = synth.object_points()
    n_tracked = n_bits = n_above = 0
    worst = 0.0
    for c0 in range(0, n, chunk):
        traj = np.array([synth.trajectory(3000 + i, 2) for i in range(c0, c0 + chunk)])
        pa = _render(ctx1080, cam, traj[:, 0], np.arange(c0, c0 + chunk))
        pb = _render(ctx1080, cam, traj[:, 1], np.arange(c0, c0 + chunk) + 1)
        pts = np.stack([synth.project(obj, traj[i, 0], cam) for i in range(chunk)]).astype(np.float32)
        out, st, err = [t.cpu().numpy() for t in ctx1080.lk(pa, pb, pts)]
        fa, fb = pa.frames.cpu().numpy(), pb.frames.cpu().numpy()
        for i in range(chunk):
            ro, rs, re = lk_oracle.lk_cv(fa[i], fb[i], pts[i])
            assert np.array_equal(st[i], rs), f"pair {c0 + i}: status differs at {np.nonzero(st[i] != rs)[0]}"
            m = rs == 1
            d = np.abs(out[i][m] - ro[m]).max(axis=1)
            n_tracked += int(m.sum())
            n_above += int((d > util.FLOW_TOL).sum())
            bad = (_bits(out[i]) != _bits(ro)).any(axis=1) & m
            n_bits += int(bad.sum())
            if bad.any():          # keep the evidence: frames, points and both answers of the pair
                print(f"pair {c0 + i}: corners {np.nonzero(bad)[0]} differ: gpu {out[i][bad]} cv {ro[bad]}")
                dump = util.ROOT / "gpurun_out"
                if dump.is_dir():
                    np.savez_compressed(dump / f"lk_mismatch_pair{c0 + i}.npz", prev=fa[i], next=fb[i], pts=pts[i], gpu=out[i], cv=ro,
                                        gpu_err=err[i], cv_err=re, status=rs)
            worst = max(worst, float(d.max()) if d.size else 0.0)
            assert np.array_equal(_bits(err[i][m]), _bits(re[m])), f"pair {c0 + i}: error differs"
        del pa, pb
    print(f"lk config-3 sweep: {n} pairs, {n_tracked} tracked corners, worst {worst:.3e} px, above 0.01 px: {n_above}, not bit-identical: {n_bits}")
    assert n_tracked > 20000
    assert n_above == 0, f"{n_above} corners differ from OpenCV by more than 0.01 px (worst {worst})"
    assert n_bits == 0, f"{n_bits} tracked corners are not bit-identical to OpenCV (worst {worst} px)"


@pytest.mark.parametrize("levels", [4, 3, 2])
def test_lk_level_parallel_equals_warp_per_corner(lib_built, levels, monkeypatch):
    """Small batches run lk_levels_kernel (a warp per pyramid level, the levels' setups in parallel, the searches chained through
    shared memory); large ones lk_kernel (a warp per corner).  Same code per level: points, status and error must be bit-identical,
    with and without the fallback rule (frames with >= 2 tags skipped), corners outside the frame and NaNs included."""
    from accurate_aprilgroup_tracking_b200.context import AgtContext
    cam = synth.CAMERA_VGA
    n = 5
    traj = synth.trajectory(3300, n + 1)
    obj = synth.object_points()
    pts = np.stack([synth.project(obj, traj[i], cam) for i in range(n)]).astype(np.float32)
    pts[0, 0] = [-40, -40]; pts[1, 1] = [np.nan, 3]; pts[2, 2] = [639.7, 479.2]; pts[3, 3] = [2000, 100]
    n_tags = np.array([0, 2, 1, 5, 0], np.int32)
    got = {}
    for label, limit in (("levels", "100000"), ("corner", "0")):
        monkeypatch.setenv("AGT_LK_SPLIT_MAX", limit)
        ctx = AgtContext(0, cam.mtx, None)
        prev = _render(ctx, cam, traj[:-1], np.arange(n) + 3300, levels=levels)
        nxt = _render(ctx, cam, traj[1:], np.arange(n) + 3301, levels=levels)
        a = [x.cpu().numpy() for x in ctx.lk(prev, nxt, pts)]
        b = [x.cpu().numpy() for x in ctx.lk(prev, nxt, pts, n_tags=ctx._dev(n_tags, ctx.torch.int32))]
        got[label] = a + b
        ctx.close()
    for x, y in zip(got["levels"], got["corner"]):
        assert np.array_equal(x.view(np.uint8), y.view(np.uint8))
    assert got["levels"][1].sum() > 100 and (got["levels"][4][[1, 3]] == 0).all()


def test_lk_textureless_all_lost(ctxvga):
    pyr_a = ctxvga.alloc_pyramid(1, 640, 480, 4)
    pyr_b = ctxvga.alloc_pyramid(1, 640, 480, 4)
    pyr_a.levels[0].fill_(128)
    pyr_b.levels[0].fill_(128)
    ctxvga.build_pyramid(pyr_a)
    ctxvga.build_pyramid(pyr_b)
    pts = np.array([[[100, 100], [320, 240], [600, 400]]], np.float32)
    out, st, err = ctxvga.lk(pyr_a, pyr_b, pts)
    assert int(st.sum()) == 0          # minEig gate: same as OpenCV on a flat image


def test_lk_empty_batch(ctxvga):
    pyr = ctxvga.alloc_pyramid(1, 640, 480, 4)
    out, st, err = ctxvga.lk(pyr, pyr, np.zeros((1, 0, 2), np.float32))
    assert out.shape == (1, 0, 2)


@pytest.mark.parametrize("cam_name,max_flow,extra", [("1080p", 32, False), ("1080p", 0, False), ("vga", 32, True), ("vga", 2, False)])
def test_lk_on_roi_pyramids_is_exact(ctx1080, ctxvga, cam_name, max_flow, extra):
    """agt_lk_roi (pyramid of the new frames built only around the corners, frames that looked outside redone on complete
    pyramids) returns bit for bit what agt_lk returns on complete pyramids - also when the rectangles are too small
    (max_flow 0 / 2: the exactness net does the work) and with corners on and outside the image border.  Levels 1-3 of the
    ROI pyramids are poisoned first, so anything read outside what was built would show."""
    import torch
    ctx, cam = (ctx1080, synth.CAMERA_1080P) if cam_name == "1080p" else (ctxvga, synth.CAMERA_VGA)
    n = 12
    trajs = [synth.trajectory(3300 + i, 6) for i in range(n)]
    first = np.array([t[0] for t in trajs])
    second = np.array([t[5 if max_flow < 32 else 1] for t in trajs])        # five frames apart: flows beyond a small max_flow
    prev = _render(ctx, cam, first, np.arange(n) + 3300)
    full = _render(ctx, cam, second, np.arange(n) + 3400)
    obj = synth.object_points()
    pts = np.stack([synth.project(obj, p, cam) for p in first]).astype(np.float32)
    if extra:
        more = np.array([[5, 5], [636.5, 3.2], [-3, 10], [700, 100], [320, 479.5], [0, 0], [639, 479], [-30, -30], [np.nan, 1.0]], np.float32)
        pts = np.concatenate([pts, np.broadcast_to(more, (n,) + more.shape)], axis=1).astype(np.float32)
    ref_out, ref_st, ref_err = ctx.lk(prev, full, pts)
    roi = ctx.alloc_pyramid(n, cam.width, cam.height, 4)
    roi.levels[0].copy_(full.levels[0])
    for l in range(1, 4):
        roi.levels[l].fill_(0xA5)
    # max_flow 0: the rectangles are computed from the first corner alone, so most other corners look outside them
    only_first = None
    if max_flow == 0:
        only_first = np.zeros(pts.shape[:2], np.uint8)
        only_first[:, 0] = 1
    out, st, err, rects, redo = ctx.lk_roi(prev, roi, pts, max_flow=max_flow, valid=only_first)
    assert torch.equal(st, ref_st)
    assert torch.equal(out.view(torch.int32), ref_out.view(torch.int32)), "tracked points differ from the complete-pyramid path"
    assert torch.equal(err.view(torch.int32), ref_err.view(torch.int32))
    n_redo = int(redo.sum())
    if max_flow >= 32 and not extra:
        print("lk roi: frames redone", n_redo, "of", n)
        assert n_redo <= n // 4, f"{n_redo} of {n} frames left a rectangle sized for 32 px of flow"
        r = rects.cpu().numpy()
        assert ((r[:, 2] - r[:, 0]) * (r[:, 3] - r[:, 1])).mean() < 0.5 * cam.width * cam.height or cam_name == "vga"
    if max_flow == 0:
        assert n_redo > n // 2, "rectangles around one corner cannot cover the other 47"
    # the previous pyramid may be a region-of-interest one as well (it is the `next` of the step before)
    prev_roi = ctx.alloc_pyramid(n, cam.width, cam.height, 4)
    prev_roi.levels[0].copy_(prev.levels[0])
    for l in range(1, 4):
        prev_roi.levels[l].fill_(0x5A)
    rp = ctx.lk_rects(prev_roi, pts, None, max_flow)
    ctx.build_pyramid_roi(prev_roi, rp)
    for l in range(1, 4):
        roi.levels[l].fill_(0xA5)
    out2, st2, err2, _, redo2 = ctx.lk_roi(prev_roi, roi, pts, rects_prev=rp, max_flow=max_flow)
    assert torch.equal(st2, ref_st) and torch.equal(out2.view(torch.int32), ref_out.view(torch.int32))
    assert torch.equal(err2.view(torch.int32), ref_err.view(torch.int32))


# ------------------------------------------------------------------------------------------
# K3 PnP vs cv2.solvePnP(SOLVEPNP_ITERATIVE)
# ------------------------------------------------------------------------------------------
def _pnp_inputs(cam, n, seed, noise=0.1):
    rng = np.random.default_rng(seed)
    obj = synth.object_points().astype(np.float32)
    img = np.zeros((n, 48, 2), np.float32)
    valid = np.zeros((n, 48), np.uint8)
    poses = np.zeros((n, 6))
    for i in range(n):
        while True:
            p = synth.random_pose(rng)
            vis = synth.visible_tags(p)
            if len(vis) >= 2:
                break
        poses[i] = p
        uv = synth.project(obj.astype(np.float64), p, cam) + rng.normal(0, noise, (48, 2))
        img[i] = uv
        for k in vis:
            valid[i, 4 * k:4 * k + 4] = 1
    return obj, img, valid, poses


@pytest.mark.parametrize("with_guess", [False, True])
def test_pnp_matches_opencv(ctxvga, with_guess):
    import cv2
    cam = synth.CAMERA_VGA
    n = 200
    obj, img, valid, truth = _pnp_inputs(cam, n, 42 + with_guess)
    rng = np.random.default_rng(7)
    guess = truth + np.concatenate([rng.normal(0, 0.03, (n, 3)), rng.normal(0, 0.003, (n, 3))], axis=1)
    pose, ok, err, iters = ctxvga.pnp(obj, img, valid, guess if with_guess else None,
                                      np.ones(n, np.uint8) if with_guess else None)
    pose, ok, err, iters = pose.cpu().numpy(), ok.cpu().numpy(), err.cpu().numpy(), iters.cpu().numpy()
    assert ok.all()
    from oracle import ape_oracle
    for i in range(n):
        m = valid[i] == 1
        if with_guess:
            r0, t0 = guess[i, :3].reshape(3, 1).copy(), guess[i, 3:].reshape(3, 1).copy()
            okr, r, t = cv2.solvePnP(obj[m], img[i][m], cam.mtx, None, r0, t0, True, flags=cv2.SOLVEPNP_ITERATIVE)
        else:
            okr, r, t = cv2.solvePnP(obj[m], img[i][m], cam.mtx, None, flags=cv2.SOLVEPNP_ITERATIVE)
        assert okr
        util.assert_pose_close(pose[i], np.concatenate([r.ravel(), t.ravel()]), f"frame {i}")
        e_ref = ape_oracle.mean_reprojection_error(obj[m], img[i][m], r, t, cam.mtx, None)
        assert abs(err[i] - e_ref) < 1e-3
    print("pnp iters mean", iters.mean(), "max", iters.max())


def test_pnp_with_distortion(ctxvga):
    import cv2
    from accurate_aprilgroup_tracking_b200.context import AgtContext
    cam = synth.CAMERA_VGA
    dist = np.array([[-0.21, 0.09, 0.0012, -0.0008, -0.015]])
    ctx = AgtContext(0, cam.mtx, dist)
    rng = np.random.default_rng(5)
    obj = synth.object_points().astype(np.float32)
    n = 32
    img = np.zeros((n, 48, 2), np.float32)
    valid = np.zeros((n, 48), np.uint8)
    for i in range(n):
        p = synth.random_pose(rng)
        vis = synth.visible_tags(p)
        while len(vis) < 2:
            p = synth.random_pose(rng)
            vis = synth.visible_tags(p)
        uv, _ = cv2.projectPoints(obj.astype(np.float64), p[:3], p[3:], cam.mtx, dist)
        img[i] = uv.reshape(-1, 2) + rng.normal(0, 0.1, (48, 2))
        for k in vis:
            valid[i, 4 * k:4 * k + 4] = 1
    pose, ok, err, _ = ctx.pnp(obj, img, valid)
    pose = pose.cpu().numpy()
    proj = ctx.project(obj, pose).cpu().numpy()
    for i in range(n):
        m = valid[i] == 1
        okr, r, t = cv2.solvePnP(obj[m], img[i][m], cam.mtx, dist, flags=cv2.SOLVEPNP_ITERATIVE)
        util.assert_pose_close(pose[i], np.concatenate([r.ravel(), t.ravel()]), f"frame {i}")
        ref, _ = cv2.projectPoints(obj.astype(np.float64), pose[i, :3], pose[i, 3:], cam.mtx, dist)
        assert np.abs(proj[i] - ref.reshape(-1, 2)).max() < 1e-6
    ctx.close()


def test_pnp_too_few_points_not_ok(ctxvga):
    cam = synth.CAMERA_VGA
    obj = synth.object_points().astype(np.float32)
    pose = synth.trajectory(60, 1)[0]
    proj = synth.project(obj, pose, cam).astype(np.float32)
    img = np.stack([np.zeros((48, 2), np.float32), np.zeros((48, 2), np.float32), proj])
    valid = np.zeros((3, 48), np.uint8)
    valid[0, :4] = 1                      # four points that all project to one pixel: no homography
    valid[2, [0, 1, 2, 3, 4]] = 1         # five points, not coplanar: the DLT needs >= 6 (cv2.solvePnP raises there)
    pose_out, ok, err, _ = ctxvga.pnp(obj, img, valid)
    assert ok.cpu().numpy().tolist() == [0, 0, 0]


def test_pnp_planar_points_start_from_a_homography(ctxvga):
    """cv.solvePnP(ITERATIVE) without a guess on coplanar object points - a single tag (4 corners), a planar tag board - takes
    the homography branch of cvFindExtrinsicCameraParams2; so does K3 (round 1 ran the 12x12 DLT on a rank-deficient system).
    Against cv2 on the same points, with and without lens distortion."""
    import cv2
    cam = synth.CAMERA_VGA
    rng = np.random.default_rng(71)
    s = synth.TAG_SIZE / 2
    tag = np.array([[-s, -s, 0], [-s, s, 0], [s, s, 0], [s, -s, 0]])
    board = np.concatenate([tag + np.array([dx, dy, 0.0]) for dx in (-0.03, 0.0, 0.03) for dy in (-0.03, 0.0, 0.03)])       # 36 points
    tilt = synth.rodrigues(np.array([0.4, -0.3, 0.2]))
    sets = [tag, board, board @ tilt.T + np.array([0.01, -0.02, 0.005])]       # the third: a plane that is not z = 0
    dist = np.array([[-0.21, 0.08, 0.0006, -0.0004, -0.01]])
    n_cases = 0
    for d in (None, dist):
        c = ctxvga if d is None else None
        if d is not None:
            from accurate_aprilgroup_tracking_b200.context import AgtContext
            c = AgtContext(0, cam.mtx, d)
        for pts in sets:
            n = len(pts)
            B = 24
            poses = np.array([synth.random_pose(rng) for _ in range(B)])
            poses[:, :3] *= 0.7                                       # keep the plane well inside the field of view
            poses[:, 5] = rng.uniform(0.15, 0.35, B)
            obj = np.zeros((48, 3), np.float32) if n <= 48 else None
            obj[:n] = pts
            img = np.zeros((B, 48, 2), np.float32)
            valid = np.zeros((B, 48), np.uint8)
            valid[:, :n] = 1
            for b in range(B):
                uv = cv2.projectPoints(pts, poses[b, :3], poses[b, 3:], cam.mtx, d)[0].reshape(-1, 2)
                img[b, :n] = uv + rng.normal(0, 0.05, uv.shape)
            got, ok, err, _ = c.pnp(obj, img, valid)
            got, ok = got.cpu().numpy(), ok.cpu().numpy()
            for b in range(B):
                okc, rc, tc = cv2.solvePnP(pts.astype(np.float32), img[b, :n], cam.mtx, d, flags=cv2.SOLVEPNP_ITERATIVE)
                assert okc and ok[b] == 1, (n, b)
                util.assert_pose_close(got[b], np.concatenate([rc.ravel(), tc.ravel()]), f"{n} coplanar points, frame {b}, dist {d is not None}")
                n_cases += 1
        if d is not None:
            c.close()
    assert n_cases == 2 * 3 * 24


def test_reprojection_gate_near_the_threshold(ctxvga):
    """A3 at the 2 px boundary (detect_pose.py:539 with transform_helper.py:98-121): one corner per frame is pushed off by a
    ramp of offsets so that the mean reprojection error of the solved pose sweeps through 2.0 px; the device error must agree
    with cv2.solvePnP + the reference's error formula to 1e-4 px there, and the gate must decide as the reference does on
    every frame whose reference error is further than that from the threshold."""
    import cv2
    from oracle import ape_oracle
    cam = synth.CAMERA_VGA
    n = 240
    obj, img, valid, truth = _pnp_inputs(cam, n, 4242, noise=0.05)
    ntags = (valid.reshape(n, 12, 4).sum(axis=2) == 4).sum(axis=1).astype(np.int32)
    rng = np.random.default_rng(11)
    guess = truth + np.concatenate([rng.normal(0, 0.01, (n, 3)), rng.normal(0, 0.001, (n, 3))], axis=1)

    def ref_error(i, pts):
        m = valid[i] == 1
        r0, t0 = guess[i, :3].reshape(3, 1).copy(), guess[i, 3:].reshape(3, 1).copy()
        okr, r, t = cv2.solvePnP(obj[m], pts[m], cam.mtx, None, r0, t0, True, flags=cv2.SOLVEPNP_ITERATIVE)
        assert okr
        return ape_oracle.mean_reprojection_error(obj[m], pts[m], r, t, cam.mtx, None)

    for i in range(n):
        # offset of one corner, found by bisection on the reference, that leaves a mean error of 2.0 +- 0.1 px (a ramp over the batch)
        target = ape_oracle.MAX_MEAN_ERROR + (i - n / 2 + 0.5) / n * 0.2
        m0 = np.flatnonzero(valid[i])[0]
        ang = rng.uniform(0, 2 * np.pi)
        u = np.array([np.cos(ang), np.sin(ang)], np.float32)
        lo, hi = 0.0, 200.0
        for _ in range(30):
            mid = 0.5 * (lo + hi)
            pts = img[i].copy()
            pts[m0] += np.float32(mid) * u
            if ref_error(i, pts) < target:
                lo = mid
            else:
                hi = mid
        img[i, m0] += np.float32(0.5 * (lo + hi)) * u
    pose, ok, err, _ = ctxvga.pnp(obj, img, valid, guess, np.ones(n, np.uint8))
    gate = ctxvga.accept_gate(ok, err, ctxvga._dev(ntags, ctxvga.torch.int32)).cpu().numpy()
    ok, err = ok.cpu().numpy(), err.cpu().numpy()
    assert ok.all()
    e_ref = np.array([ref_error(i, img[i]) for i in range(n)])
    tol = 1e-4
    assert np.abs(err - e_ref).max() < tol, np.abs(err - e_ref).max()
    near = np.abs(e_ref - ape_oracle.MAX_MEAN_ERROR) < 0.05
    assert near.sum() >= 10 and (e_ref[near] < 2.0).any() and (e_ref[near] > 2.0).any(), "the sweep missed the threshold"
    decided = np.abs(e_ref - ape_oracle.MAX_MEAN_ERROR) > tol
    want = (e_ref < ape_oracle.MAX_MEAN_ERROR) & (ntags >= ape_oracle.MIN_TAGS)
    assert np.array_equal(gate[decided] != 0, want[decided])
    print("gate sweep: errors", e_ref.min(), "..", e_ref.max(), "within 0.05 px of the threshold:", int(near.sum()),
          "max |err - ref|", np.abs(err - e_ref).max())


# ------------------------------------------------------------------------------------------
# K3 + K0: the APE state machine over whole streams vs the reference restatement
# ------------------------------------------------------------------------------------------
def test_ape_streams_match_reference_state_machine(ctxvga):
    from oracle import ape_oracle
    cam = synth.CAMERA_VGA
    torch = ctxvga.torch
    n_streams, n_frames = 6, 60
    group = ape_oracle.group_from_json(synth.april_group_dict())
    oracles = [ape_oracle.ApeOracle(group, cam.mtx, None, True) for _ in range(n_streams)]
    trajs = [synth.trajectory(5000 + s, n_frames) for s in range(n_streams)]
    rngs = [np.random.default_rng(5000 + s) for s in range(n_streams)]
    obj = synth.object_points().astype(np.float32)
    state = ctxvga.new_stream_state(n_streams)
    for f in range(n_frames):
        img = np.zeros((n_streams, 48, 2), np.float32)
        valid = np.zeros((n_streams, 48), np.uint8)
        ntags = np.zeros(n_streams, np.int32)
        all_dets = []
        for s in range(n_streams):
            dets = synth.detections(trajs[s][f], cam, rngs[s])
            if s == 1 and f in (20, 21):
                dets = dets[:1]                                   # tracking loss: < 2 tags
            if s == 2 and f == 30:
                dets = [(t, c + (25.0 if i == 0 else 0.0)) for i, (t, c) in enumerate(dets)]   # gate failure
            all_dets.append(dets)
            ntags[s] = len(dets)
            for t, c in dets:
                img[s, 4 * t:4 * t + 4] = c
                valid[s, 4 * t:4 * t + 4] = 1
        guess, use = ctxvga.ape_prepare(state)
        pose, ok, err, _ = ctxvga.pnp(obj, img, valid, guess, use)
        acc, flag = ctxvga.ape_update(state, ntags, pose, ok, err)
        st = state.cpu().numpy()
        acc = acc.cpu().numpy()
        for s in range(n_streams):
            oracles[s].step(all_dets[s])
            snap = oracles[s].snapshot()
            assert bool(acc[s]) == oracles[s].last_accepted, (f, s)
            assert (st[s, 7] != 0) == (snap["guess"] is not None), (f, s)
            assert (st[s, 0] != 0) == (snap["prev"] is not None), (f, s)
            if snap["prev"] is not None:
                util.assert_pose_close(st[s, 1:7], np.concatenate(snap["prev"]), f"prev f{f} s{s}")
            if snap["guess"] is not None:
                util.assert_pose_close(st[s, 8:14], np.concatenate(snap["guess"]), f"guess f{f} s{s}")
            assert int(st[s, 16]) == snap["n_vel"]


# ------------------------------------------------------------------------------------------
# K4 dense refinement vs oracle/dpr_oracle.py
# ------------------------------------------------------------------------------------------
def _dpr_case(ctx, cam, n, seed, sig_r=0.01, sig_t=0.0005, n_hyp=1):
    from oracle import dpr_oracle
    rng = np.random.default_rng(seed)
    truth = np.array([synth.random_pose(rng) for _ in range(n)])
    pyr = _render(ctx, cam, truth, np.arange(n) + seed)
    init = truth[:, None, :] + np.concatenate([rng.normal(0, sig_r, (n, n_hyp, 3)), rng.normal(0, sig_t, (n, n_hyp, 3))], axis=2)
    res = ctx.refine(pyr, init, n_hyp)
    out = {k: v.cpu().numpy() for k, v in res.items()}
    model = util.dpr_model()
    levels = [[pyr.level(l)[b].cpu().numpy() for l in range(4)] for b in range(n)]
    return truth, init, out, model, levels, res, pyr


def test_dpr_matches_oracle_1080p(ctx1080):
    from oracle import dpr_oracle
    cam = synth.CAMERA_1080P
    n = 24
    truth, init, out, model, levels, _, _ = _dpr_case(ctx1080, cam, n, 2000)
    worst_r = worst_t = 0.0
    ev_gpu, ev_ref = [], []
    for b in range(n):
        ref = dpr_oracle.refine(levels[b], model, cam.mtx, init[b, 0])
        dr, dt = util.pose_diff(out["pose"][b, 0], ref["pose"])
        worst_r, worst_t = max(worst_r, dr), max(worst_t, dt)
        ev_gpu.append(int(out["evals"][b, 0])); ev_ref.append(ref["evals"])
        assert out["n_valid"][b, 0] == ref["n_valid"], (b, out["n_valid"][b, 0], ref["n_valid"])
        assert out["status"][b, 0] == ref["status"], (b, out["status"][b, 0], ref["status"])
        util.assert_pose_close(out["pose"][b, 0], ref["pose"], f"frame {b}")
        assert abs(out["cost"][b, 0] - ref["cost"]) <= 1e-3 * ref["cost"]
    print("dpr worst", worst_r, worst_t, "evals gpu", ev_gpu, "ref", ev_ref)


def test_dpr_matches_oracle_vga(ctxvga):
    from oracle import dpr_oracle
    cam = synth.CAMERA_VGA
    n = 12
    truth, init, out, model, levels, _, _ = _dpr_case(ctxvga, cam, n, 2100, 0.005, 0.0003)
    for b in range(n):
        ref = dpr_oracle.refine(levels[b], model, cam.mtx, init[b, 0])
        util.assert_pose_close(out["pose"][b, 0], ref["pose"], f"frame {b}")
        assert out["n_valid"][b, 0] == ref["n_valid"]


@pytest.mark.parametrize("cam_name", ["vga", "1080p"])
def test_dpr_kernel_lands_on_scipys_fixed_point(ctxvga, ctx1080, cam_name):
    """The 64 + 64 pin frames of tests/golden/dpr_pin.npz (noisy renders, BASELINE config 2's perturbation): the kernel's pose
    against the pose scipy found with its own solvers (oracle/dpr_pin.py) - kernel == independent solver, within a tenth of
    the parity bar - and against the oracle's committed answer."""
    from oracle import dpr_pin
    pin = np.load(dpr_pin.GOLDEN)
    ctx = ctxvga if cam_name == "vga" else ctx1080
    cam = dpr_pin.CAMERAS[cam_name]
    sel = np.nonzero(pin["cam"] == (0 if cam_name == "vga" else 1))[0]
    frames, inits = [], []
    for k in sel:
        truth, init, frame = dpr_pin.case(cam_name, int(pin["index"][k]))
        frames.append(frame); inits.append(init)
    pyr = ctx.alloc_pyramid(len(sel), cam.width, cam.height, 4)
    ctx.upload_frames(pyr, np.stack(frames))
    ctx.build_pyramid(pyr)
    out = {k: v.cpu().numpy() for k, v in ctx.refine(pyr, np.stack(inits).reshape(len(sel), 1, 6), 1).items()}
    worst = np.zeros(4)
    for j, k in enumerate(sel):
        assert out["status"][j, 0] == 1
        ds = util.pose_diff(out["pose"][j, 0], pin["scipy_pose"][k])
        do = util.pose_diff(out["pose"][j, 0], pin["oracle_pose"][k])
        worst = np.maximum(worst, [ds[0], ds[1], do[0], do[1]])
        assert ds[0] <= 1e-5 and ds[1] <= 2e-6, f"{cam_name} case {pin['index'][k]}: kernel is {ds[0]:.2e} rad, {ds[1]:.2e} m from scipy's fixed point"
        util.assert_pose_close(out["pose"][j, 0], pin["oracle_pose"][k], f"{cam_name} case {pin['index'][k]} vs oracle")
    ev = out["evals"][:, 0]
    print(f"{cam_name}: kernel vs scipy worst {worst[0]:.2e} rad {worst[1]:.2e} m; vs oracle {worst[2]:.2e} rad {worst[3]:.2e} m; "
          f"evaluations mean {ev.mean():.2f} (oracle {pin['oracle_evals'][sel].mean():.2f}), identical counts {int((ev == pin['oracle_evals'][sel]).sum())}/{len(sel)}")


def test_dpr_multi_hypothesis_selection(ctx1080):
    """BASELINE config 4 at its shape: 64 hypotheses per 1080p frame, truth + N(0, 0.03 rad), N(0, 2 mm); the index
    agt_select_best returns is the index oracle/dpr_oracle.py:refine_multi returns, and so is the pose."""
    from oracle import dpr_oracle
    cam = synth.CAMERA_1080P
    n, n_hyp = 3, 64
    truth, init, out, model, levels, res, _ = _dpr_case(ctx1080, cam, n, 2200, 0.03, 0.002, n_hyp)
    best, best_pose = ctx1080.select_best(res)
    best, best_pose = best.cpu().numpy(), best_pose.cpu().numpy()
    for b in range(n):
        want, runs = dpr_oracle.refine_multi(levels[b], model, cam.mtx, init[b])
        score = np.array([2.0 * r["cost"] / r["n_valid"] if r["n_valid"] > 0 else np.inf for r in runs])
        gpu_score = np.where(out["n_valid"][b] > 0, 2.0 * out["cost"][b].astype(np.float64) / np.maximum(out["n_valid"][b], 1), np.inf)
        n_tied = int((score <= score.min() * (1 + dpr_oracle.SELECT_TIE)).sum())
        print(f"frame {b}: oracle winner {want}, kernel winner {best[b]}, {n_tied} of {n_hyp} runs within the tie band, "
              f"score spread {score.min():.6f} .. {score.max():.6f}")
        assert best[b] == want, f"frame {b}: kernel picked {best[b]}, oracle {want} (scores {gpu_score[best[b]]}, {score[want]})"
        assert np.array_equal(best_pose[b], out["pose"][b, best[b]])
        util.assert_pose_close(best_pose[b], runs[want]["pose"], f"frame {b} winner")
        # every hypothesis of the frame against the oracle run from the same start
        for h in range(n_hyp):
            if runs[h]["status"] == dpr_oracle.ST_CONVERGED and out["status"][b, h] == 1:
                util.assert_pose_close(out["pose"][b, h], runs[h]["pose"], f"frame {b} hypothesis {h}")


def test_dpr_multi_hypothesis_exact_ties_take_the_lowest_index(ctx1080):
    cam = synth.CAMERA_1080P
    truth, init, out, model, levels, res, pyr = _dpr_case(ctx1080, cam, 2, 2250, 0.01, 0.0005, 1)
    dup = np.repeat(init, 5, axis=1)
    dup[:, 0] = truth[:, None, :][:, 0] + np.array([0.2, 0.2, 0.2, 0.01, 0.01, 0.02])        # a bad start first: must not win
    res = ctx1080.refine(pyr, dup, 5)
    best, _ = ctx1080.select_best(res)
    assert best.cpu().numpy().tolist() == [1, 1]


def test_dpr_fixed_point_idempotent(ctx1080):
    # size-independent property: refining an already refined pose (same visible set, converged)
    # moves it by less than the parity tolerance
    cam = synth.CAMERA_1080P
    truth, init, out, model, levels, res, pyr = _dpr_case(ctx1080, cam, 16, 2300)
    again = {k: v.cpu().numpy() for k, v in ctx1080.refine(pyr, res["pose"], 1).items()}
    checked = 0
    for b in range(16):
        if out["status"][b, 0] != 1 or again["n_valid"][b, 0] != out["n_valid"][b, 0]:
            continue        # visibility / level is re-frozen at the new initial pose: a different cost function
        checked += 1
        util.assert_pose_close(again["pose"][b, 0], out["pose"][b, 0], f"frame {b}")
    assert checked >= 8


def test_dpr_object_outside_image(ctx1080):
    # pose whose projection is far outside the frame: no valid samples -> status NONE, pose unchanged
    cam = synth.CAMERA_1080P
    pyr = ctx1080.alloc_pyramid(1, cam.width, cam.height, 4)
    pyr.levels[0].fill_(128)
    ctx1080.build_pyramid(pyr)
    init = np.array([[[0.1, 0.2, 0.3, 5.0, 0.0, 0.4]]])
    res = ctx1080.refine(pyr, init, 1)
    assert int(res["status"][0, 0]) == 0 and int(res["n_valid"][0, 0]) == 0
    util.assert_pose_close(res["pose"][0, 0].cpu().numpy(), init[0, 0])


# ------------------------------------------------------------------------------------------
# region-of-interest pyramid + refinement: identical to the full-frame path
# ------------------------------------------------------------------------------------------
def test_roi_pyramid_and_refinement_are_exact(ctx1080):
    import cv2
    cam = synth.CAMERA_1080P
    torch = ctx1080.torch
    rng = np.random.default_rng(2500)
    n = 20
    truth = np.array([synth.random_pose(rng) for _ in range(n)])
    truth[0, 3:] = (0.16, 0.09, 0.27)         # object cut by the image corner
    truth[1, 3:] = (0.0, 0.0, 0.2505)         # largest ROI of the working volume
    init = truth + np.concatenate([rng.normal(0, 0.01, (n, 3)), rng.normal(0, 0.0005, (n, 3))], axis=1)
    init[2, 3] += 0.010                       # starts 10 mm off: the LM run leaves its predicted ROI -> redone on the full pyramid
    full = ctx1080.alloc_pyramid(n, cam.width, cam.height, 4)
    ctx1080.render(full, truth, np.arange(n) + 2500)
    roi = ctx1080.alloc_pyramid(n, cam.width, cam.height, 4)
    roi.levels[0].copy_(full.levels[0])
    for l in (1, 2, 3):
        roi.levels[l].fill_(255)              # poison: anything the ROI path does not compute stays visibly wrong
    ctx1080.build_pyramid(full)
    want = ctx1080.refine(full, init.reshape(n, 1, 6), 1)
    got = ctx1080.refine_roi(roi, init.reshape(n, 1, 6), 1)
    for key in ("pose", "cost", "n_valid", "evals", "status"):
        assert torch.equal(got[key], want[key]), key
    redo = got["redo"].cpu().numpy()
    assert redo[2] == 1 and redo.sum() <= 3
    # inside the rectangle (minus the pyrDown halo) the ROI levels equal cv2.pyrDown of the frame
    rects = ctx1080.dpr_rects(roi, init.reshape(n, 1, 6), 1).cpu().numpy()
    for b in (0, 1, 5):
        x0, y0, x1, y1 = rects[b]
        assert x0 % 16 == 0 and 0 <= x0 < x1 <= cam.width and 0 <= y0 < y1 <= cam.height
        ref = full.frames[b].cpu().numpy()
        for l in (1, 2, 3):
            ref = cv2.pyrDown(ref)
            pad = (2 << l) >> l
            sl = (slice((y0 >> l) + pad + 1, (y1 >> l) - pad - 1), slice((x0 >> l) + pad + 1, (x1 >> l) - pad - 1))
            if redo[b] or sl[0].stop <= sl[0].start:
                continue
            assert np.array_equal(roi.level(l)[b].cpu().numpy()[sl], ref[sl]), (b, l)
    # the ROI path touched only a small part of the levels
    assert float((roi.level(1)[5] == 255).float().mean()) > 0.5


def test_bgr_to_gray_bit_exact(ctxvga):
    import cv2
    rng = np.random.default_rng(3)
    for (w, h) in [(640, 480), (1920, 1080), (333, 77)]:
        bgr = rng.integers(0, 256, (2, h, w, 3), dtype=np.uint8)
        pyr = ctxvga.alloc_pyramid(2, w, h, 1)
        ctxvga.ingest_bgr(pyr, bgr)
        got = pyr.frames.cpu().numpy()
        for b in range(2):
            assert np.array_equal(got[b], cv2.cvtColor(bgr[b], cv2.COLOR_BGR2GRAY)), (w, h, b)


# ------------------------------------------------------------------------------------------
# K1 fused into K4
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n_hyp", [1, 3])
def test_fused_pyramid_refinement_is_exact(ctx1080, n_hyp):
    # K1 fused into K4: only level 0 given, every refinement builds its own level ROI; identical to pyramid + refine.
    # Depths chosen to hit every pyramid level (q = fx * pitch / z: level 0 above 0.42 m, 1, 2, and 3 below 0.105 m),
    # objects cut by the image border (reflected taps), a large batch (one CTA per pose) and small ones (clusters).
    cam = synth.CAMERA_1080P
    torch = ctx1080.torch
    rng = np.random.default_rng(2600 + n_hyp)
    for n in (40, 3):
        truth = np.array([synth.random_pose(rng) for _ in range(n)])
        truth[0, 3:] = (0.16, 0.09, 0.27)          # cut by the image corner, level 1
        truth[1, 3:] = (-0.20, -0.11, 0.30)        # cut by the opposite corner
        truth[2, 3:] = (0.01, -0.02, 0.16)         # level 2
        if n > 3:
            truth[3, 3:] = (0.0, 0.0, 0.50)        # level 0: nothing to build
            truth[4, 3:] = (0.005, 0.004, 0.10)    # level 3
            truth[5, 3:] = (0.06, 0.03, 0.12)      # level 2/3 boundary, cut by the border
            truth[6, 3:] = (0.0, 0.0, 0.2505)
        init = truth[:, None, :] + np.concatenate([rng.normal(0, 0.01, (n, n_hyp, 3)), rng.normal(0, 0.0005, (n, n_hyp, 3))], axis=2)
        full = ctx1080.alloc_pyramid(n, cam.width, cam.height, 4)
        ctx1080.render(full, truth, np.arange(n) + 2600)
        base = ctx1080.alloc_pyramid(n, cam.width, cam.height, 4)
        base.levels[0].copy_(full.levels[0])
        for l in (1, 2, 3):
            base.levels[l].fill_(255)             # poison: a refinement that reads what it did not build goes visibly wrong
        ctx1080.build_pyramid(full)
        want = ctx1080.refine(full, init, n_hyp)
        got = ctx1080.refine(base, init, n_hyp, fused=True)
        inside = (got["left_roi"] == 0)
        assert float(inside.float().mean()) > 0.9
        assert torch.equal(got["left_roi"], want["left_roi"])
        for key in ("pose", "cost", "n_valid", "evals", "status"):
            a, b = got[key][inside], want[key][inside]
            assert torch.equal(a, b), (key, n, n_hyp)
        levels_used = set()
        for b in range(n):
            q = cam.mtx[0, 0] * synth.model_pitch() / init[b, 0, 5]
            levels_used.add(0 if q < 2 else 1 if q < 4 else 2 if q < 8 else 3)
        if n > 3:
            assert levels_used == {0, 1, 2, 3}
            # what was built equals the full pyramid there, and most of every level was never touched
            for b, l in ((0, 1), (2, 2), (4, 3)):
                built = base.level(l)[b] != 255
                assert bool(built.any())
                assert torch.equal(base.level(l)[b][built], full.level(l)[b][built])
            assert float((base.level(1)[3] == 255).float().mean()) == 1.0      # level-0 refinement builds nothing
        # the exactness net (redo of frames that left their ROI) gives the full-pyramid result everywhere
        got2 = ctx1080.refine_fused(base, init, n_hyp)
        for key in ("pose", "cost", "n_valid", "evals", "status"):
            assert torch.equal(got2[key], want[key]), (key, "after redo")


@pytest.mark.parametrize("w,h,f", [(637, 479, 600.0), (333, 201, 300.0), (161, 121, 150.0)])
def test_fused_pyramid_odd_sizes(lib_built, w, h, f):
    # odd widths (unaligned rows: no 16-byte copies), odd heights and levels only a few pixels wide
    from accurate_aprilgroup_tracking_b200.context import AgtContext
    cam = synth.Camera(w, h, f, f, w / 2.0, h / 2.0)
    ctx = AgtContext(0, cam.mtx, None)
    ctx.set_synthetic_model()
    try:
        torch = ctx.torch
        rng = np.random.default_rng(w)
        pitch = synth.model_pitch()
        depths = [f * pitch / q for q in (1.5, 3.0, 6.0, 8.5)]           # one pose per pyramid level (the last: lens almost on the tag)
        n = len(depths)
        truth = np.array([synth.random_pose(rng) for _ in range(n)])
        for b, z in enumerate(depths):
            truth[b, 3:] = (0.2 * z * (b - 1.5) / 1.5 * w / (2 * f), 0.0, z)
        init = truth + np.concatenate([rng.normal(0, 0.005, (n, 3)), rng.normal(0, 0.0002, (n, 3))], axis=1)
        full = ctx.alloc_pyramid(n, w, h, 4)
        ctx.render(full, truth, np.arange(n) + w)
        base = ctx.alloc_pyramid(n, w, h, 4)
        base.levels[0].copy_(full.levels[0])
        for l in (1, 2, 3):
            base.levels[l].fill_(255)
        ctx.build_pyramid(full)
        want = ctx.refine(full, init.reshape(n, 1, 6), 1)
        got = ctx.refine_fused(base, init.reshape(n, 1, 6), 1)
        assert int(want["n_valid"][0]) > 0          # (short focal lengths put the deeper-level poses inside the object)
        for key in ("pose", "cost", "n_valid", "evals", "status"):
            assert torch.equal(got[key], want[key]), (key, w, h)
    finally:
        ctx.close()


# ------------------------------------------------------------------------------------------
# frame ingest with undistortion (row N2): cv.undistort + crop + BGR2GRAY in one pass, bit exact
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("w,h,f", [(640, 480, 600.0), (1920, 1080, 1400.0), (333, 201, 300.0)])
def test_undistort_ingest_bit_exact(lib_built, w, h, f):
    import cv2
    from accurate_aprilgroup_tracking_b200.context import AgtContext
    from accurate_aprilgroup_tracking_b200.cv_compat import HostContext
    rng = np.random.default_rng(h)
    mtx = np.array([[f, 0, w / 2 + 3.3], [0, f * 1.01, h / 2 - 2.1], [0, 0, 1]])
    for dist in ([-0.28, 0.11, 0.0007, -0.0004, -0.02], [0.05, -0.1, 0.001, 0.002, 0.03]):
        dist = np.array(dist, np.float64).reshape(1, 5)
        new_mtx, roi = cv2.getOptimalNewCameraMatrix(mtx, dist, (w, h), 1, (w, h))
        x, y, rw, rh = roi
        ctx = AgtContext(0, mtx, dist)
        try:
            ctx.set_undistort(new_mtx, w, h, roi)
            # 11 / 37 frames: a partial group and three groups (the last one partial) of the per-CTA frame loop
            for n, shape in ((11, (h, w, 3)), (3, (h, w))) + (((37, (h, w, 3)),) if w == 640 else ()):
                frames = rng.integers(0, 256, (n,) + shape, dtype=np.uint8)
                frames[0] = cv2.GaussianBlur(frames[0], (0, 0), 3.0)
                pyr = ctx.alloc_pyramid(n, rw, rh, 1)
                ctx.ingest_undistort(pyr, frames)
                got = pyr.frames.cpu().numpy()
                for b in range(n):
                    want = cv2.undistort(frames[b], mtx, dist, None, new_mtx)[y:y + rh, x:x + rw]
                    want = cv2.cvtColor(want, cv2.COLOR_BGR2GRAY) if want.ndim == 3 else want
                    assert np.array_equal(got[b], want), (w, h, b, int(np.abs(got[b].astype(int) - want.astype(int)).max()))
            # a new camera matrix that zooms out three times: the source box of a 32x32 tile no longer fits a shared-memory
            # stage, so the tiles take the global-memory path; and one that zooms in (boxes of a few pixels)
            for zoom in (1.0 / 3.0, 2.5):
                zk = np.array([[f * zoom, 0, w / 2], [0, f * zoom, h / 2], [0, 0, 1]])
                ctx.set_undistort(zk, w, h, (0, 0, w, h))
                frames = rng.integers(0, 256, (2, h, w, 3), dtype=np.uint8)
                pyr = ctx.alloc_pyramid(2, w, h, 1)
                ctx.ingest_undistort(pyr, frames)
                got = pyr.frames.cpu().numpy()
                for b in range(2):
                    want = cv2.cvtColor(cv2.undistort(frames[b], mtx, dist, None, zk), cv2.COLOR_BGR2GRAY)
                    assert np.array_equal(got[b], want), (w, h, zoom, b)
        finally:
            ctx.close()
        host = HostContext(0)
        frame = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        want = cv2.cvtColor(cv2.undistort(frame, mtx, dist, None, new_mtx)[y:y + rh, x:x + rw], cv2.COLOR_BGR2GRAY)
        assert np.array_equal(host.undistort_gray(frame, mtx, dist, new_mtx, roi), want)
        host.close()



def test_undistort_ingest_large_batch(lib_built):
    """Large batches put more frames on a CTA (16 below 128 frames, 32 below 256, 64 from there on: fewer pixel-map
    evaluations): same bits as cv2 and whatever the grouping."""
    import cv2
    from accurate_aprilgroup_tracking_b200.context import AgtContext
    w, h, f = 320, 200, 280.0
    rng = np.random.default_rng(77)
    mtx = np.array([[f, 0, w / 2 - 1.7], [0, f, h / 2 + 0.9], [0, 0, 1]])
    dist = np.array([[-0.31, 0.12, 0.001, -0.0007, -0.03]])
    new_mtx, roi = cv2.getOptimalNewCameraMatrix(mtx, dist, (w, h), 1, (w, h))
    x, y, rw, rh = roi
    ctx = AgtContext(0, mtx, dist)
    try:
        ctx.set_undistort(new_mtx, w, h, roi)
        n = 300                                                    # 4 full groups of 64 + one of 44
        frames = rng.integers(0, 256, (n, h, w, 3), dtype=np.uint8)
        big = ctx.alloc_pyramid(n, rw, rh, 1)
        ctx.ingest_undistort(big, frames)
        got = big.frames.cpu().numpy()
        for m in (100, 150):                                       # groups of 16 / of 32
            part = ctx.alloc_pyramid(m, rw, rh, 1)
            ctx.ingest_undistort(part, frames[:m])
            assert np.array_equal(part.frames.cpu().numpy(), got[:m]), m
        for b in (0, 63, 64, 255, 256, 299):
            want = cv2.cvtColor(cv2.undistort(frames[b], mtx, dist, None, new_mtx)[y:y + rh, x:x + rw], cv2.COLOR_BGR2GRAY)
            assert np.array_equal(got[b], want), b
    finally:
        ctx.close()


def _exact_window(lo, hi, n0, n, level):
    """Exact part [a, b) of a level-`level` axis of a pyramid built under the level-0 interval [lo, hi) (agt.h)."""
    reach = (2 << level) - 2
    a = 0 if lo <= 0 else (lo + reach + (1 << level) - 1) >> level
    b = n if hi >= n0 else ((hi - 1 - reach) >> level) + 1
    return a, max(b, 0)


def test_roi_pyramid_single_launch_chain(ctxvga):
    """Batches that fill the machine build the region-of-interest pyramid of every frame in one launch (a CTA per frame
    walks down the levels): inside the window of pixels whose pyrDown support lies in the frame's rectangle it equals
    the complete pyramid bit for bit, outside the windows K1 would write it leaves the (poisoned) levels untouched;
    empty rectangles, rectangles on the image border and full-frame rectangles included."""
    import torch
    ctx = ctxvga
    w, h, n = 1280, 128, 320          # more than 2 frames per SM: the single-launch path; 16 px per lane on levels 1-2, 8 on level 3
    rng = np.random.default_rng(11)
    frames = torch.as_tensor(rng.integers(0, 256, (n, h, w), dtype=np.uint8), device=ctx.tdev)
    full = ctx.alloc_pyramid(n, w, h, 4)
    full.levels[0].copy_(frames)
    ctx.build_pyramid(full)
    rects = np.zeros((n, 4), np.int32)
    for b in range(n):
        x0, y0 = int(rng.integers(0, w - 48)) & ~15, int(rng.integers(0, h - 40))
        x1, y1 = min(w, (x0 + int(rng.integers(40, 900)) + 15) & ~15), min(h, y0 + int(rng.integers(30, 128)))
        rects[b] = (x0, y0, x1, y1)
    rects[0] = (0, 0, w, h)                                            # whole frame
    rects[1] = (0, 0, 0, 0)                                            # empty: nothing is built
    rects[2] = (0, 0, 64, 50)                                          # top-left corner
    rects[3] = (w - 64, h - 50, w, h)                                  # bottom-right corner
    roi = ctx.alloc_pyramid(n, w, h, 4)
    roi.levels[0].copy_(frames)
    for l in (1, 2, 3):
        roi.levels[l].fill_(0xA5)
    ctx.build_pyramid_roi(roi, torch.as_tensor(rects, device=ctx.tdev))
    small = ctx.alloc_pyramid(24, w, h, 4)                             # the per-level launches on a small batch
    small.levels[0].copy_(frames[:24])
    for l in (1, 2, 3):
        small.levels[l].fill_(0xA5)
    ctx.build_pyramid_roi(small, torch.as_tensor(rects[:24], device=ctx.tdev))
    for l in (1, 2, 3):
        got, want, per_level = roi.level(l).cpu().numpy(), full.level(l).cpu().numpy(), small.level(l).cpu().numpy()
        # (the two paths may start a window on different 8 / 16 pixel boundaries: compare them where both are exact)
        lw, lh = w >> l, h >> l
        for b in range(n):
            x0, y0, x1, y1 = rects[b]
            if x1 <= x0 or y1 <= y0:
                assert (got[b] == 0xA5).all()
                continue
            ax, bx = _exact_window(x0, x1, w, lw, l)
            ay, by = _exact_window(y0, y1, h, lh, l)
            if bx > ax and by > ay:
                assert np.array_equal(got[b, ay:by, ax:bx], want[b, ay:by, ax:bx]), (l, b)
                if b < 24:
                    assert np.array_equal(per_level[b, ay:by, ax:bx], want[b, ay:by, ax:bx]), (l, b)
            # nothing is written outside the largest window K1 may touch (16-pixel alignment on the left)
            sh, rnd = l, (1 << l) - 1
            mx0, my0 = max(0, (x0 >> sh) - 2) & ~15, max(0, (y0 >> sh) - 2)
            mx1, my1 = min(lw, (((x1 + rnd) >> sh) + 2 + 7) & ~7), min(lh, ((y1 + rnd) >> sh) + 2)
            outside = np.ones_like(got[b], bool)
            outside[my0:my1, mx0:mx1] = False
            assert (got[b][outside] == 0xA5).all(), (l, b)
    assert np.array_equal(roi.level(1)[0].cpu().numpy(), full.level(1)[0].cpu().numpy())


def test_undistort_ingest_random_cameras(lib_built):
    """Bit-exactness of the undistort ingest over random lenses, new camera matrices (alpha 0..1, off-centre, skew-free),
    crops and 16-byte aligned frame sizes, BGR and gray."""
    import cv2
    from accurate_aprilgroup_tracking_b200.context import AgtContext
    rng = np.random.default_rng(2024)
    checked = 0
    for case in range(14):
        w = int(rng.choice([64, 176, 320, 640, 1008]))
        h = int(rng.integers(33, 300))
        f = float(rng.uniform(0.5, 1.6)) * w
        mtx = np.array([[f, 0, w / 2 + rng.uniform(-9, 9)], [0, f * rng.uniform(0.97, 1.03), h / 2 + rng.uniform(-9, 9)], [0, 0, 1]])
        dist = np.array([[rng.uniform(-0.35, 0.2), rng.uniform(-0.15, 0.15), rng.uniform(-0.004, 0.004), rng.uniform(-0.004, 0.004),
                          rng.uniform(-0.05, 0.05)]])
        new_mtx, roi = cv2.getOptimalNewCameraMatrix(mtx, dist, (w, h), float(rng.uniform(0, 1)), (w, h))
        x, y, rw, rh = roi
        if rw < 8 or rh < 8:
            continue
        ctx = AgtContext(0, mtx, dist)
        try:
            ctx.set_undistort(new_mtx, w, h, roi)
            for shape in ((5, h, w, 3), (2, h, w)):
                frames = rng.integers(0, 256, shape, dtype=np.uint8)
                pyr = ctx.alloc_pyramid(shape[0], rw, rh, 1)
                ctx.ingest_undistort(pyr, frames)
                got = pyr.frames.cpu().numpy()
                for b in range(shape[0]):
                    want = cv2.undistort(frames[b], mtx, dist, None, new_mtx)[y:y + rh, x:x + rw]
                    want = cv2.cvtColor(want, cv2.COLOR_BGR2GRAY) if want.ndim == 3 else want
                    assert np.array_equal(got[b], want), (case, w, h, roi, b, int(np.abs(got[b].astype(int) - want.astype(int)).max()))
            checked += 1
        finally:
            ctx.close()
    assert checked >= 8
