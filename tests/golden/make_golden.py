"""Generate tests/golden/*.npz by IMPORTING the parts of the reference that run in the authoring container
(utils/transformation.py, utils/calibration.py, utils/preprocess.py) and recording their
outputs on seeded inputs. /root/reference does not exist on the GPU box, so the vectors are committed.

    python tests/golden/make_golden.py        # needs /root/reference
"""
import os
import sys
import types

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def import_reference_utils():
    for name in ("ipdb", "turtle"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.pos = None
            m.set_trace = lambda *a, **k: None
            sys.modules[name] = m
    sys.path.insert(0, REF)
    from utils import transformation, preprocess  # noqa
    from utils import calibration  # noqa
    return transformation, calibration, preprocess


def main():
    T, Cal, P = import_reference_utils()
    rng = np.random.default_rng(13)
    out = {}

    # ---- Kabsch: get_rigid_transform_3D (utils/transformation.py:178-222)
    refs, tgts, Rs, ts, ns = [], [], [], [], []
    kmax = 10
    for case in range(64):
        k = int(rng.integers(3, kmax + 1))
        a = rng.normal(0, 0.08, (k, 3))
        if case % 8 == 0:
            a[:, 2] = 0.0                      # planar key points
        q = rng.normal(size=4)
        q /= np.linalg.norm(q)
        Rg = T.get_quaternion_rotation_matrix(q, switch_w=False)
        b = a @ Rg.T + rng.normal(0, 1.0, 3) + rng.normal(0, 0.002, (k, 3))
        if case % 16 == 5:
            b = b * np.array([1, 1, -1.0])     # mirrored target: exercises the reflection branch
        R, t = T.get_rigid_transform_3D(a, b)
        pa, pb = np.zeros((kmax, 3)), np.zeros((kmax, 3))
        pa[:k], pb[:k] = a, b
        refs.append(pa); tgts.append(pb); Rs.append(R); ts.append(t); ns.append(k)
    out.update(kabsch_ref=np.array(refs), kabsch_tgt=np.array(tgts), kabsch_R=np.array(Rs), kabsch_t=np.array(ts),
               kabsch_n=np.array(ns, dtype=np.int32))

    # ---- quaternion / pose algebra
    quats = rng.normal(size=(32, 4))
    quats /= np.linalg.norm(quats, axis=1, keepdims=True)
    poses = np.concatenate((rng.normal(0, 1, (32, 3)), quats), axis=1)
    poses2 = np.concatenate((rng.normal(0, 1, (32, 3)), np.roll(quats, 5, axis=0)), axis=1)
    out.update(
        quat=quats,
        quat_matrix=np.array([T.get_quaternion_rotation_matrix(q, switch_w=False) for q in quats]),
        quat_matrix_switch=np.array([T.get_quaternion_rotation_matrix(np.roll(q, -1), switch_w=True) for q in quats]),
        pose=poses, pose2=poses2,
        pose_matrix=np.array([T.get_transformation_matrix(p, switch_w=False) for p in poses]),
        pose_roundtrip=np.array([T.get_pose_from_matrix(T.get_transformation_matrix(p)) for p in poses]),
        pose_inverse=np.array([T.get_pose_inverse(p) for p in poses]),
        base2cam=np.array([T.get_base2cam_pose(p, p2) for p, p2 in zip(poses, poses2)]),
        pose2pose=np.array([T.transform_pose2pose(p, p2) for p, p2 in zip(poses, poses2)]),
        switch_w=np.array([T.switch_w(np.concatenate((p[:3], np.roll(p[3:], -1)))) for p in poses]),
    )

    # ---- calibration averaging (utils/calibration.py:69-139)
    base = poses[0]
    noisy = np.tile(base, (10, 1))
    noisy[:, :3] += rng.normal(0, 0.01, (10, 3))
    noisy[:, 3:] += rng.normal(0, 0.01, (10, 4))
    noisy[:, 3:] /= np.linalg.norm(noisy[:, 3:], axis=1, keepdims=True)
    w = rng.random(10) + 0.5
    out.update(avg_in=noisy, avg_w=w, avg_out=Cal.compute_poses_average(noisy),
               avg_out_w=Cal.compute_poses_average(noisy, weights=w),
               outlier_in=np.concatenate((rng.normal(0, 1, 20), [15.0, -12.0])),
               )
    out["outlier_flags"] = Cal.get_outliers(out["outlier_in"])[0]

    # ---- preprocess (utils/preprocess.py:8-37)
    pts = rng.normal(0, 0.2, (500, 3)).astype(np.float32)
    rgb255 = (rng.random((500, 3)) * 255).astype(np.float32)
    rgb01 = rng.random((500, 3)).astype(np.float32)
    c, off = P.center_at_origin(pts)
    out.update(pre_pts=pts, pre_centered=c, pre_offset=off, pre_rgb255=rgb255, pre_rgb255_out=P.normalize_colors(rgb255),
               pre_rgb01=rgb01, pre_rgb01_out=P.normalize_colors(rgb01))
    # the "negative input" branch (per-channel min-max rescale, utils/preprocess.py:28-32) and already-centred input;
    # own generator so that the entries recorded before these were added keep their values
    rng2 = np.random.default_rng(99)
    rgbneg = (rng2.random((400, 3)) * 1.5 - 0.3).astype(np.float32)
    rgbcen = (rng2.random((400, 3)) - 0.5).astype(np.float32)
    rgbneg255 = (rng2.random((400, 3)) * 300 - 20).astype(np.float32)
    out.update(pre_rgbneg=rgbneg, pre_rgbneg_out=P.normalize_colors(rgbneg), pre_rgbcen=rgbcen,
               pre_rgbcen_out=P.normalize_colors(rgbcen), pre_rgbneg255=rgbneg255,
               pre_rgbneg255_out=P.normalize_colors(rgbneg255))
    # ---- per-point heads: the reference's own utils/output.py (:45-87) and utils/metrics.py (:110-127), imported with
    #      this repo's oracle package standing in for `MinkowskiEngine` (output.py only uses it in a type annotation)
    import torch
    sys.path.insert(0, os.path.join(HERE, "..", ".."))
    import oracle.MinkowskiEngine as OME
    sys.modules["MinkowskiEngine"] = OME
    from utils import output as RO
    from utils import metrics as RM
    g = torch.Generator().manual_seed(13)
    kp_logits = torch.randn(700, 6, generator=g) * 6.0
    kp_logits[:, 3] = kp_logits[:, 3] * 0.05 - 3.0    # a class that never gets confident
    idx75, cls75, pr75 = RO.get_key_point_predictions(kp_logits, 0.75)
    idx999, cls999, pr999 = RO.get_key_point_predictions(kp_logits, 0.999)
    kp10_logits = torch.randn(900, 10, generator=g) * 9.0
    idx10, cls10, pr10 = RO.get_key_point_predictions(kp10_logits, 0.75)
    vote_out = torch.randn(1200, 2, generator=g)
    vote_coords = rng.normal(0, 0.1, (1200, 3)).astype(np.float32)
    vote_q = quats[3].astype(np.float32)
    vc_plain = RO.get_pred_center(vote_out, vote_coords.copy(), ee_r=0.02, q=None)
    vc_q = RO.get_pred_center(vote_out, vote_coords.copy(), ee_r=0.02, q=vote_q)
    seg_logits = torch.randn(5000, 3, generator=g)
    seg_logits[::7, 1] = seg_logits[::7, 0]           # exact ties: torch.max returns the lowest index

    class _Field:
        features = seg_logits
    seg_preds, seg_conf = RO.get_segmentations_from_tensor_field(_Field())
    pm = [RM.compute_pose_metrics(a.copy(), b.copy()) for a, b in zip(poses, poses2)]
    out.update(kp_logits=kp_logits.numpy(), kp_idx75=idx75, kp_cls75=cls75, kp_pr75=pr75.numpy(), kp_idx999=idx999,
               kp_cls999=cls999, kp_pr999=pr999.numpy(), kp10_logits=kp10_logits.numpy(), kp10_idx=idx10,
               kp10_cls=cls10, kp10_pr=pr10.numpy(), vote_out=vote_out.numpy(), vote_coords=vote_coords,
               vote_q=vote_q, vote_center=np.asarray(vc_plain), vote_center_q=np.asarray(vc_q),
               seg_logits=seg_logits.numpy(), seg_preds=seg_preds, seg_conf=seg_conf,
               metric_dist=np.array([m["dist_position"] for m in pm]),
               metric_angle=np.array([m["angle_diff"] for m in pm]))
    # ---- ROI mask (utils/data.py:58-75), reference's own function
    from utils import data as RD
    roi_pts = rng.normal(0, 1.0, (4000, 3)).astype(np.float32)
    roi_pts[::97] = 0.75          # points exactly on a limit: strict inequalities drop them
    roi_args = dict(min_x=-0.8, max_x=0.75, min_y=-1.2, max_y=0.9, min_z=-0.75, max_z=2.0)
    out.update(roi_pts=roi_pts, roi_limits=np.array([roi_args[k] for k in ("min_x", "max_x", "min_y", "max_y", "min_z",
                                                                             "max_z")], dtype=np.float64),
               roi_mask=RD.get_roi_mask(roi_pts, **roi_args), roi_mask_default=RD.get_roi_mask(roi_pts),
               roi_mask_offset=RD.get_roi_mask(roi_pts, offset=0.1, **roi_args))

    # ---- "magic" translation: the reference's own InferenceEngine.predict_translation (app/inference_engine.py:459-489),
    #      imported with stub modules for what is not installed (open3d, tensorboardX, openpyxl: none is touched by
    #      this method) and the oracle package as MinkowskiEngine; utils/config.py parses sys.argv at import time
    import tempfile
    import types
    tmp = tempfile.mkdtemp(prefix="b2me_golden_")
    argv0 = sys.argv
    sys.argv = ["x", "--config", os.path.join(REF, "config", "default.yaml"), "--log_path", os.path.join(tmp, "log.log"),
                "--exp_path", os.path.join(tmp, "exp")]
    for name in ("open3d", "tensorboardX", "openpyxl"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.SummaryWriter = object
            sys.modules[name] = m
    import oracle.MinkowskiEngine.modules.resnet_block as _rb
    import oracle.MinkowskiEngine.utils as _mu
    import oracle.MinkowskiEngine.MinkowskiOps as _mo
    sys.modules["MinkowskiEngine.modules"] = OME.modules
    sys.modules["MinkowskiEngine.modules.resnet_block"] = _rb
    sys.modules["MinkowskiEngine.utils"] = _mu
    sys.modules["MinkowskiEngine.MinkowskiOps"] = _mo
    sys.path.insert(0, os.path.join(REF, "app"))
    from app import inference_engine as IE
    sys.argv = argv0
    eng = object.__new__(IE.InferenceEngine)     # predict_translation only reads the module-level config
    tr_pts, tr_q, tr_out, tr_n = [], [], [], []
    nmax = 3000
    for case in range(8):
        n = int(rng.integers(600, nmax + 1))
        qq = rng.normal(size=4)
        qq /= np.linalg.norm(qq)
        Rq = T.get_quaternion_rotation_matrix(qq, switch_w=False)
        box = (rng.random((n, 3)) * np.array([0.10, 0.22, 0.126]) - np.array([0.0, 0.11, 0.0]))
        pts = (box @ Rq.T + np.array([rng.uniform(-0.3, 0.3), rng.uniform(-0.2, 0.2), rng.uniform(1.0, 1.5)])
               ).astype(np.float32)
        pos, _ = eng.predict_translation(pts, None, qq.astype(np.float32))
        pad = np.zeros((nmax, 3), np.float32)
        pad[:n] = pts
        tr_pts.append(pad); tr_q.append(qq.astype(np.float32)); tr_out.append(np.asarray(pos, dtype=np.float64))
        tr_n.append(n)
    out.update(trans_pts=np.array(tr_pts), trans_q=np.array(tr_q), trans_out=np.array(tr_out),
               trans_n=np.array(tr_n, dtype=np.int32))
    # ---- calibration tail: the reference's own InferenceEngine.calibrate (app/inference_engine.py:152-244) on seeded
    #      per-frame results of 3 robot positions (some frames unconfident, some without a key-point pose)
    from dto import ResultDTO
    eng.camera_link_transformation_pose = np.array([0.0, -0.045, 0.0, 0.5, -0.5, 0.5, 0.5], dtype=np.float32)
    cal_rows = []
    data = {}
    true_base = poses[5].copy()
    for pos_id in range(3):
        lst = []
        for fr in range(6):
            def noisy(p, s):
                o = p.copy()
                o[:3] += rng.normal(0, s, 3)
                o[3:] += rng.normal(0, s, 4)
                o[3:] /= np.linalg.norm(o[3:])
                return o
            ee = noisy(poses[pos_id], 0.01)
            conf = not (fr == 1 and pos_id == 0)
            has_kp = not (fr == 2)
            d = ResultDTO(segmentation=None, ee_pose=ee, base_pose=noisy(true_base, 0.01),
                          key_points_pose=noisy(poses[pos_id], 0.01) if has_kp else None,
                          key_points_base_pose=noisy(true_base, 0.01) if has_kp else None, is_confident=conf)
            lst.append(d)
            nanp = np.full(7, np.nan)
            cal_rows.append(np.concatenate(([pos_id, float(conf)], d.ee_pose, d.base_pose,
                                            d.key_points_pose if has_kp else nanp,
                                            d.key_points_base_pose if has_kp else nanp)))
        data[str(pos_id)] = lst
    import io
    import contextlib
    with contextlib.redirect_stdout(io.StringIO()):
        cal_all = eng.calibrate(data)
        cal_one = eng.calibrate({"0": data["0"]})
    out.update(calib_rows=np.array(cal_rows), calib_camera_link=eng.camera_link_transformation_pose,
               calib_pose=np.asarray(cal_all.pose_camera_link), calib_base=np.asarray(cal_all.base_pose),
               calib_kp_base=np.asarray(cal_all.key_points_base_pose),
               calib_base_cl=np.asarray(cal_all.base_pose_camera_link),
               calib_one_pose=np.asarray(cal_one.pose_camera_link),
               calib_one_base_cl=np.asarray(cal_one.base_pose_camera_link))
    for k in [k for k in sys.modules if k.startswith("MinkowskiEngine")]:
        sys.modules.pop(k)
    np.savez_compressed(os.path.join(HERE, "reference_geometry.npz"), **out)

    # ---- CAD input fixture: xyz of app/hand_files/hand.pcd (4480 points), read with our PCD reader
    sys.modules.pop("MinkowskiEngine", None)   # the alias above must not shadow the CUDA package b200calib imports
    sys.path.insert(0, os.path.join(HERE, "..", "..", "markerless-robot-camera-calibration_b200"))
    from b200calib.icp import read_pcd_xyz
    cad = read_pcd_xyz(os.path.join(REF, "app", "hand_files", "hand.pcd"))
    np.savez_compressed(os.path.join(HERE, "cad_hand_points.npz"), xyz=cad.astype(np.float32))
    print("kabsch cases", len(ns), "cad points", cad.shape, cad.min(0), cad.max(0))


if __name__ == "__main__":
    main()
