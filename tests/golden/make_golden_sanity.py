"""Generate tests/golden/reference_sanity.npz from the reference's OWN per-frame sanity check:
utils/data.py get_6_key_points (:255-335), utils/metrics.py compute_kp_error (:130-136) and
InferenceEngine.check_sanity (app/inference_engine.py:246-279), run in the authoring container on seeded EE clouds
built from the reference's CAD points (tests/golden/cad_hand_points.npz = xyz of app/hand_files/hand.pcd).

Patched: stub modules for what is not installed (ipdb, open3d, tensorboardX, openpyxl, turtle), the oracle package as
MinkowskiEngine (import only), np.int / np.long aliases (removed in NumPy 2, used at utils/data.py:273 and
app/inference_engine.py:269).

    python tests/golden/make_golden_sanity.py        # needs /root/reference
"""
import os
import sys
import tempfile
import types

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
NMAX = 4480


def main():
    tmp = tempfile.mkdtemp(prefix="b2me_golden_")
    sys.argv = ["x", "--config", os.path.join(REF, "config", "default.yaml"), "--log_path", os.path.join(tmp, "log.log"),
                "--exp_path", os.path.join(tmp, "exp")]
    for name in ("ipdb", "open3d", "tensorboardX", "openpyxl", "turtle"):
        m = types.ModuleType(name)
        m.set_trace = lambda *a, **k: None
        m.SummaryWriter = object
        m.pos = None
        sys.modules[name] = m
    for alias in ("int", "long"):
        if not hasattr(np, alias):
            setattr(np, alias, int)
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "markerless-robot-camera-calibration_b200"))
    import oracle.MinkowskiEngine as OME
    import oracle.MinkowskiEngine.modules.resnet_block as rb
    import oracle.MinkowskiEngine.utils as mu
    import oracle.MinkowskiEngine.MinkowskiOps as mo
    sys.modules["MinkowskiEngine"] = OME
    sys.modules["MinkowskiEngine.modules"] = OME.modules
    sys.modules["MinkowskiEngine.modules.resnet_block"] = rb
    sys.modules["MinkowskiEngine.utils"] = mu
    sys.modules["MinkowskiEngine.MinkowskiOps"] = mo
    sys.path.insert(0, REF)
    sys.path.insert(0, os.path.join(REF, "app"))
    from app import inference_engine as IE
    from dto import PointCloudDTO, ResultDTO
    from utils import data as RD
    from utils import metrics as RM
    from utils.transformation import get_quaternion_rotation_matrix

    cad = np.load(os.path.join(HERE, "cad_hand_points.npz"))["xyz"].astype(np.float64)
    rng = np.random.default_rng(29)
    pts_all = np.zeros((24, NMAX, 3), np.float32)
    n_all = np.zeros(24, np.int32)
    pose_all = np.zeros((24, 7), np.float64)           # x y z qw qx qy qz
    kp_w = np.zeros((24, 6, 3), np.float64)            # get_6_key_points(switch_w=False, threshold 0.04)
    idx_w = np.zeros((24, 6), np.int64)
    kp_x = np.zeros((24, 6, 3), np.float64)            # get_6_key_points(switch_w=True, default threshold) on xyzw poses
    idx_x = np.zeros((24, 6), np.int64)
    pred_cls = np.full((24, 6), -1, np.int64)
    pred_xyz = np.zeros((24, 6, 3), np.float32)
    n_pred = np.zeros(24, np.int32)
    kp_err = np.zeros(24, np.float64)
    sane = np.zeros(24, bool)
    empty_w = np.zeros(24, bool)
    empty_x = np.zeros(24, bool)
    seg_all = np.zeros((24, NMAX), np.int64)
    for case in range(24):
        q = rng.normal(size=4)
        q /= np.linalg.norm(q)
        R = get_quaternion_rotation_matrix(q, switch_w=False)
        t = np.array([rng.uniform(-0.5, 0.5), rng.uniform(-0.5, 0.5), rng.uniform(0.8, 1.5)])
        keep = rng.random(len(cad)) < rng.uniform(0.55, 1.0)
        mode = case % 6
        if mode == 1:
            keep &= ~((cad[:, 2] > 0.08) & (cad[:, 1] > 0))      # left gripper side hidden: mirrored from the right
        elif mode == 2:
            keep &= ~((cad[:, 2] > 0.08) & (cad[:, 1] < 0))
        elif mode == 3:
            keep &= cad[:, 1] < 0.06                               # a corner region missing: template corner kept
        elif mode == 4:
            keep &= rng.random(len(cad)) < 0.35                    # sparse: below min_num_of_ee_points
        local = cad[keep] + rng.normal(0, 0.0016, (int(keep.sum()), 3))
        cam = (local @ R.T + t).astype(np.float32)
        n = len(cam)
        # predicted pose = true pose, slightly off in some cases (corners then miss their templates)
        dq = rng.normal(0, 0.02 if mode != 5 else 0.15, 4)
        qp = q + dq
        qp /= np.linalg.norm(qp)
        pose = np.concatenate((t + rng.normal(0, 0.003 if mode != 5 else 0.03, 3), qp))
        # background + arm points around it; labels 0 / 1 / 2
        bg = (rng.random((NMAX - n, 3)) * 2 - 1).astype(np.float32) + np.array([0, 0, 2.5], np.float32)
        points = np.concatenate((cam, bg))
        seg = np.concatenate((np.full(n, 2), rng.integers(0, 2, NMAX - n)))
        perm = rng.permutation(NMAX)
        points, seg = points[perm], seg[perm]
        ee = points[seg == 2]
        k1, i1 = RD.get_6_key_points(ee, pose, switch_w=False, euclidean_threshold=0.04)
        pose_xyzw = np.concatenate((pose[:3], pose[4:], pose[3:4]))
        k2, i2 = RD.get_6_key_points(ee, pose_xyzw)
        # predicted key points: some of the gt key points plus noise (classes in random order), or too few of them
        m = int(rng.integers(2, 7))
        cls = rng.permutation(6)[:m]
        base = k1[cls] if len(k1) else np.tile(pose[:3], (m, 1))   # empty: no EE point in front of the pose
        xyz = (base + rng.normal(0, 0.01 if case % 3 else 0.06, (m, 3))).astype(np.float32)
        res = ResultDTO(segmentation=seg, ee_pose=pose, key_points=[(int(c), x) for c, x in zip(cls, xyz)])
        ok = IE.InferenceEngine.check_sanity(None, PointCloudDTO(points=points, rgb=np.zeros_like(points), timestamp=None), res)
        pts_all[case], n_all[case], pose_all[case], seg_all[case] = points, n, pose, seg
        empty_w[case], empty_x[case] = len(k1) == 0, len(k2) == 0
        if len(k1):
            kp_w[case], idx_w[case] = k1, i1
        if len(k2):
            kp_x[case], idx_x[case] = k2, i2
        pred_cls[case, :m], pred_xyz[case, :m], n_pred[case] = cls, xyz, m
        kp_err[case] = RM.compute_kp_error(k1, xyz, cls)
        sane[case] = ok
    print("sane:", sane.astype(int), "ee points:", n_all)
    print("corner found (switch_w=False):", (idx_w[:, :4] >= 0).sum(1))
    np.savez_compressed(os.path.join(HERE, "reference_sanity.npz"), points=pts_all, seg=seg_all.astype(np.int8),
                        n_ee=n_all, pose=pose_all, kp_w=kp_w, idx_w=idx_w, kp_x=kp_x, idx_x=idx_x, pred_cls=pred_cls,
                        pred_xyz=pred_xyz, n_pred=n_pred, kp_err=kp_err, sane=sane, empty_w=empty_w, empty_x=empty_x)


if __name__ == "__main__":
    main()
