"""GPU tests of the edge cases of the C ABI: empty and ragged inputs, maximum sizes, NaNs, argument errors."""
import ctypes as C

import numpy as np
import pytest

from accurate_aprilgroup_tracking_b200 import _lib, synth

pytestmark = pytest.mark.gpu


def test_empty_batches_are_noops(ctxvga):
    torch = ctxvga.torch
    cam = synth.CAMERA_VGA
    obj = synth.object_points().astype(np.float32)
    pose, ok, err, it = ctxvga.pnp(obj, np.zeros((0, 48, 2), np.float32), np.zeros((0, 48), np.uint8))
    assert pose.shape == (0, 6) and ok.numel() == 0
    pyr = ctxvga.alloc_pyramid(0, cam.width, cam.height, 4)
    ctxvga.build_pyramid(pyr)
    res = ctxvga.refine(pyr, np.zeros((0, 1, 6)), 1)
    assert res["pose"].shape == (0, 1, 6)
    out, st, e = ctxvga.lk(pyr, pyr, np.zeros((0, 48, 2), np.float32))
    assert out.shape == (0, 48, 2)
    assert ctxvga.project(obj, np.zeros((0, 6))).shape == (0, 48, 2)


def test_argument_errors_map_to_exceptions(ctxvga, lib_built):
    from accurate_aprilgroup_tracking_b200.context import AgtContext
    cam = synth.CAMERA_VGA
    with pytest.raises(ValueError):
        ctxvga.pnp(np.zeros((65, 3), np.float32), np.zeros((1, 65, 2), np.float32))      # > AGT_MAX_POINTS
    with pytest.raises(ValueError):
        ctxvga.set_camera(cam.mtx, np.zeros(3))                                            # 0, 4 or 5 coefficients
    with pytest.raises(ValueError):
        ctxvga.set_camera(np.diag([0.0, 600.0, 1.0]), None)                                # non-positive focal length
    s, tg, n, c = synth.surface_model()
    with pytest.raises(ValueError):
        ctxvga.set_model(s, tg[::-1].copy(), n, c, synth.model_pitch())                    # not tag-major
    with pytest.raises(ValueError):
        ctxvga.set_model(s, tg, n, c, 0.0)
    fresh = AgtContext(0)
    pyr = fresh.alloc_pyramid(1, 64, 48, 2)
    with pytest.raises(RuntimeError):
        fresh.pnp(np.zeros((8, 3), np.float32), np.zeros((1, 8, 2), np.float32))           # camera not set
    fresh.set_camera(cam.mtx, None)
    with pytest.raises(RuntimeError):
        fresh.refine(pyr, np.zeros((1, 1, 6)), 1)                                          # model not set
    fresh.set_synthetic_model()
    fresh.set_camera(cam.mtx, np.array([0.1, 0, 0, 0, 0.0]))
    with pytest.raises(ValueError):
        fresh.refine(pyr, np.zeros((1, 1, 6)), 1)                                          # refinement needs undistorted frames
    bad = _lib.AgtPyramid()
    bad.levels = 9
    with pytest.raises(ValueError):
        fresh._check(fresh.lib.agt_build_pyramid(fresh.h, C.byref(bad), 1))
    fresh.close()


def test_maximum_point_count_and_ragged_validity(ctxvga):
    """16 tags (64 corners, the ABI maximum) and frames with different subsets of valid corners."""
    import cv2
    cam = synth.CAMERA_VGA
    rng = np.random.default_rng(8)
    obj = np.concatenate([synth.object_points(), synth.object_points()[:16] * 0.97]).astype(np.float32)     # 64 points
    n = 8
    img = np.zeros((n, 64, 2), np.float32)
    valid = np.zeros((n, 64), np.uint8)
    poses = []
    for i in range(n):
        p = synth.random_pose(rng)
        poses.append(p)
        img[i] = synth.project(obj.astype(np.float64), p, cam) + rng.normal(0, 0.05, (64, 2))
        k = rng.integers(8, 65)
        valid[i, rng.permutation(64)[:k]] = 1
    valid[0, :] = 1
    pose, ok, err, _ = ctxvga.pnp(obj, img, valid)
    pose, ok = pose.cpu().numpy(), ok.cpu().numpy()
    assert ok.all()
    for i in range(n):
        m = valid[i] == 1
        _, r, t = cv2.solvePnP(obj[m], img[i][m], cam.mtx, None, flags=cv2.SOLVEPNP_ITERATIVE)
        from tests import util
        util.assert_pose_close(pose[i], np.concatenate([r.ravel(), t.ravel()]), f"frame {i} ({m.sum()} points)")


def test_nan_inputs_do_not_poison_the_batch(ctx1080):
    from tests import util
    cam = synth.CAMERA_1080P
    rng = np.random.default_rng(4)
    truth = np.array([synth.random_pose(rng) for _ in range(3)])
    pyr = ctx1080.alloc_pyramid(3, cam.width, cam.height, 4)
    ctx1080.render(pyr, truth, [1, 2, 3])
    ctx1080.build_pyramid(pyr)
    init = truth + 0.002
    clean = ctx1080.refine(pyr, init.reshape(3, 1, 6), 1)["pose"].cpu().numpy()
    init_bad = init.copy()
    init_bad[1, 4] = np.nan
    res = ctx1080.refine(pyr, init_bad.reshape(3, 1, 6), 1)
    st = res["status"].cpu().numpy().ravel()
    assert st[1] == 0 and int(res["n_valid"][1, 0]) == 0
    got = res["pose"].cpu().numpy()
    assert np.array_equal(got[0], clean[0]) and np.array_equal(got[2], clean[2])
    # LK: a NaN point is lost (status 0), the others are unaffected
    pts = np.stack([synth.project(synth.object_points(), truth[i], cam) for i in range(3)]).astype(np.float32)
    good = ctx1080.lk(pyr, pyr, pts)
    pts_bad = pts.copy()
    pts_bad[0, 5] = np.nan
    bad = ctx1080.lk(pyr, pyr, pts_bad)
    assert int(bad[1][0, 5]) == 0
    mask = np.ones((3, 48), bool); mask[0, 5] = False
    assert np.array_equal(bad[0].cpu().numpy()[mask], good[0].cpu().numpy()[mask])
    # PnP: non-finite image points are treated as absent
    obj = synth.object_points().astype(np.float32)
    img = np.stack([synth.project(obj.astype(np.float64), truth[i], cam) for i in range(3)]).astype(np.float32)
    img_bad = img.copy(); img_bad[2, 7] = np.inf
    v = np.ones((3, 48), np.uint8)
    v2 = v.copy(); v2[2, 7] = 0
    a = ctx1080.pnp(obj, img_bad, v)[0].cpu().numpy()
    b = ctx1080.pnp(obj, img, v2)[0].cpu().numpy()
    assert np.array_equal(a, b)


def test_fused_refine_and_undistort_edge_cases(ctxvga, lib_built):
    from accurate_aprilgroup_tracking_b200.context import AgtContext
    torch = ctxvga.torch
    cam = synth.CAMERA_VGA
    # empty batch, and a pose whose ROI is empty (object far outside the frame): nothing to build, status NONE
    pyr0 = ctxvga.alloc_pyramid(0, cam.width, cam.height, 4)
    assert ctxvga.refine(pyr0, np.zeros((0, 1, 6)), 1, fused=True)["pose"].shape == (0, 1, 6)
    pyr = ctxvga.alloc_pyramid(2, cam.width, cam.height, 4)
    pyr.levels[0].fill_(128)
    for l in (1, 2, 3):
        pyr.levels[l].fill_(255)
    init = np.array([[[0.1, 0.2, 0.3, 5.0, 0.0, 0.4]], [[float("nan")] * 6]])
    res = ctxvga.refine(pyr, init, 1, fused=True)
    assert res["status"].reshape(-1).tolist() == [0, 0] and res["n_valid"].reshape(-1).tolist() == [0, 0]
    assert all(bool((pyr.levels[l] == 255).all()) for l in (1, 2, 3))
    # masked frames are skipped by the fused launch too
    mask = torch.tensor([0, 1], dtype=torch.uint8, device=pyr.levels[0].device)
    out = ctxvga.refine(pyr, init, 1, mask=mask, fused=True)
    assert int(out["evals"][0, 0]) == 0
    # undistortion: must be configured first, for the frame size it is used with
    ctx = AgtContext(0, cam.mtx, np.array([-0.2, 0.05, 0.0, 0.0, 0.0]))
    try:
        g = ctx.alloc_pyramid(1, cam.width, cam.height, 1)
        with pytest.raises(RuntimeError):
            ctx.ingest_undistort(g, np.zeros((1, cam.height, cam.width, 3), np.uint8))
        with pytest.raises(ValueError):
            ctx.set_undistort(np.eye(3), cam.width, cam.height, (0, 0, cam.width + 1, cam.height))       # crop outside the frame
        with pytest.raises(ValueError):
            ctx.set_undistort(np.array([[600.0, 0, 320], [0.1, 600, 240], [0, 0, 1]]), cam.width, cam.height, (0, 0, 8, 8))
        ctx.set_undistort(cam.mtx, cam.width, cam.height, (0, 0, cam.width, cam.height))
        with pytest.raises(ValueError):
            ctx.ingest_undistort(g, np.zeros((1, cam.height // 2, cam.width, 3), np.uint8))              # another frame size
        ctx.ingest_undistort(g, np.zeros((0, cam.height, cam.width, 3), np.uint8))                        # empty batch
    finally:
        ctx.close()
