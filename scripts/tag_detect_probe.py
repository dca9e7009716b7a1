"""agt_detect_tags / agt_decode_tags against cv2.aruco and the true corners on rendered frames (run on the GPU box)."""
import sys, time, numpy as np, torch
sys.path.insert(0, '.')
from accurate_aprilgroup_tracking_b200 import synth
from accurate_aprilgroup_tracking_b200.context import AgtContext
from oracle import tag_oracle
obj = synth.object_points()
for cam in (synth.CAMERA_VGA, synth.CAMERA_1080P):
    ctx = AgtContext(0, cam.mtx, None)
    n = 24
    poses = np.array([synth.trajectory(700 + s, 1)[0] for s in range(n)])
    pyr = ctx.alloc_pyramid(n, cam.width, cam.height, 1)
    ctx.render(pyr, poses, np.arange(n))
    for win in (0, 3, 4):
        out = {k: v.cpu().numpy() for k, v in ctx.detect_tags(pyr, refine_win=win).items()}
        torch.cuda.synchronize(); t0 = time.time()
        for _ in range(5): ctx.detect_tags(pyr, refine_win=win)
        torch.cuda.synchronize(); ms = (time.time() - t0) / 5 * 1e3
        frames = pyr.frames.cpu().numpy()
        tot_vis = found = wrong = aruco_found = both = 0
        e_true, e_aruco, e_aruco_true = [], [], []
        for f in range(n):
            vis = set(synth.visible_tags(poses[f]).tolist())
            ours = {int(out["id"][f, k]): out["corners"][f, k] for k in range(out["n"][f])}
            ar = dict(tag_oracle.detect_cv(frames[f]))
            tot_vis += len(vis); found += len(set(ours) & vis); wrong += len(set(ours) - set(range(12))); aruco_found += len(set(ar) & vis)
            for i, c in ours.items():
                if i < 12:
                    true = synth.project(obj[4 * i:4 * i + 4], poses[f], cam)
                    e_true.append(np.abs(c - true).max())
                    if i in ar:
                        both += 1
                        e_aruco.append(np.abs(c - ar[i]).max()); e_aruco_true.append(np.abs(ar[i] - true).max())
            extra = set(ours) - vis
            missing = (set(ar) & vis) - set(ours)
            if missing or (set(ours) - set(range(12))):
                print(f"  frame {f}: visible {sorted(vis)} aruco {sorted(ar)} ours {sorted(ours)} missing-vs-aruco {sorted(missing)}")
        e_true, e_aruco, e_aruco_true = map(np.array, (e_true, e_aruco, e_aruco_true))
        print(f"{cam.width}x{cam.height} refine_win {win}: visible {tot_vis}, ours {found}, aruco {aruco_found}, ids outside the group {wrong}; "
              f"corner error vs truth median {np.median(e_true):.2f} max {e_true.max():.2f} px; vs aruco median {np.median(e_aruco):.2f} max {e_aruco.max():.2f}; "
              f"(aruco vs truth median {np.median(e_aruco_true):.2f} max {e_aruco_true.max():.2f}); {ms:.2f} ms per {n} frames")
    ctx.close()
