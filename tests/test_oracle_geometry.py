"""Oracle pinning (CPU): the geometry oracle and the host-side pose algebra against golden vectors produced by the
reference's own utils/transformation.py, utils/calibration.py and utils/preprocess.py (tests/golden/make_golden.py),
and against scikit-learn's single-linkage clustering (the reference's clustering back end, utils/output.py:15-28)."""
import numpy as np
import pytest

from oracle import geometry as G
from oracle import pipeline as OP


def test_oracle_kabsch_vs_reference_golden(golden):
    for i in range(len(golden["kabsch_n"])):
        k = int(golden["kabsch_n"][i])
        R, t = G.rigid_transform_3D(golden["kabsch_ref"][i][:k], golden["kabsch_tgt"][i][:k])
        assert np.allclose(R, golden["kabsch_R"][i], atol=1e-12)
        assert np.allclose(t, golden["kabsch_t"][i], atol=1e-12)


def test_oracle_pose_algebra_vs_reference_golden(golden):
    for i, q in enumerate(golden["quat"]):
        assert np.allclose(G.quaternion_rotation_matrix(q, switch_w=False), golden["quat_matrix"][i], atol=1e-15)
        assert np.allclose(G.quaternion_rotation_matrix(np.roll(q, -1), switch_w=True), golden["quat_matrix_switch"][i],
                           atol=1e-15)
    for i, p in enumerate(golden["pose"]):
        assert np.allclose(G.transformation_matrix(p), golden["pose_matrix"][i], atol=1e-15)
        assert np.allclose(G.pose_from_matrix(G.transformation_matrix(p)), golden["pose_roundtrip"][i], atol=1e-12)


def test_host_pose_algebra_vs_reference_golden(golden, built_lib):
    """b200calib.transformation mirrors utils/transformation.py (same names, same results)."""
    from b200calib import transformation as T
    for i, (p, p2) in enumerate(zip(golden["pose"], golden["pose2"])):
        assert np.allclose(T.get_transformation_matrix(p), golden["pose_matrix"][i], atol=1e-15)
        assert np.allclose(T.get_pose_inverse(p), golden["pose_inverse"][i], atol=1e-12)
        assert np.allclose(T.get_base2cam_pose(p, p2), golden["base2cam"][i], atol=1e-12)
        assert np.allclose(T.transform_pose2pose(p, p2), golden["pose2pose"][i], atol=1e-12)
        assert np.allclose(T.switch_w(np.concatenate((p[:3], np.roll(p[3:], -1)))), golden["switch_w"][i])


def test_calibration_average_vs_reference_golden(golden, built_lib):
    from b200calib import calibration as C
    a = C.compute_poses_average(golden["avg_in"])
    b = C.compute_poses_average(golden["avg_in"], weights=golden["avg_w"])
    for got, ref in ((a, golden["avg_out"]), (b, golden["avg_out_w"])):
        assert np.allclose(got[:3], ref[:3], atol=1e-12)
        assert min(np.abs(got[3:] - ref[3:]).max(), np.abs(got[3:] + ref[3:]).max()) < 1e-10  # eigenvector sign
    assert np.array_equal(C.get_outliers(golden["outlier_in"])[0], golden["outlier_flags"])


def test_preprocess_vs_reference_golden(golden):
    c, off = OP.center_at_origin(golden["pre_pts"])
    assert np.array_equal(c, golden["pre_centered"]) and np.array_equal(off, golden["pre_offset"])
    assert np.array_equal(OP.normalize_colors(golden["pre_rgb255"]), golden["pre_rgb255_out"])
    assert np.array_equal(OP.normalize_colors(golden["pre_rgb01"]), golden["pre_rgb01_out"])
    for k in ("pre_rgbneg", "pre_rgbcen", "pre_rgbneg255"):     # negative inputs: per-channel min-max branch
        assert np.array_equal(OP.normalize_colors(golden[k]), golden[k + "_out"]), k


def test_largest_cluster_vs_sklearn():
    rng = np.random.default_rng(5)
    a = rng.normal(0, 0.03, (400, 3))
    b = rng.normal(0, 0.02, (150, 3)) + [0.5, 0, 0]
    c = rng.normal(0, 0.01, (30, 3)) + [0, 0.6, 0.1]
    pts = np.concatenate((a, b, c))[rng.permutation(580)].astype(np.float32)
    assert np.array_equal(G.largest_cluster(pts, 0.06), G.largest_cluster_sklearn(pts, 0.06))
    assert len(G.largest_cluster(pts[:1], 0.06)) == 1 and len(G.largest_cluster(pts[:0], 0.06)) == 0


def test_icp_recovers_known_pose(cad_points):
    rng = np.random.default_rng(13)
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    T = G.transformation_matrix(np.concatenate(([0.1, -0.2, 1.2], q)))
    tgt = cad_points[rng.choice(len(cad_points), 3000, replace=False)].astype(np.float64) @ T[:3, :3].T + T[:3, 3]
    init = T.copy()
    init[:3, 3] += [0.01, -0.008, 0.012]
    To, fit, rmse, it = G.icp_point_to_point(cad_points, tgt.astype(np.float32), init)
    assert fit == 1.0 and np.linalg.norm(To[:3, 3] - T[:3, 3]) < 2e-3 and G.rotation_angle_deg(To[:3, :3], T[:3, :3]) < 1.0
    assert 1 <= it <= 30


def test_oracle_pipeline_small_frame(built_lib):
    """the per-frame oracle pipeline runs end to end on a small synthetic frame (gate, crops, poses, ICP)."""
    import torch
    import oracle.MinkowskiEngine as OME
    from b200calib.models import make_models, randomize_bn_stats
    from b200calib.synthetic import make_frame, ee_surface_cloud
    torch.manual_seed(3)
    M = make_models(OME)
    nets = dict(seg=randomize_bn_stats(M.RobotNetSegmentation(3, num_classes=3, variant="MinkUNet14A")).eval(),
                rot=randomize_bn_stats(M.RobotNetEncode(3, 7, variant="MinkUNet14A")).eval(),
                kp=randomize_bn_stats(M.RobotNetSegmentation(3, num_classes=6, variant="MinkUNet14A")).eval())
    f = make_frame(21, width=160, height=120)
    r = OP.predict_frame(nets, ee_surface_cloud(1024), f["points"], f["rgb"],
                         dict(seg_scale=50.0, ee_point_counts_threshold=32, kp_conf_threshold=0.0),
                         gt_labels=f["labels"])
    assert r["segmentation"].shape == (len(f["points"]),) and set(np.unique(r["segmentation"])) <= {0, 1, 2}
    assert r["ee_pose"] is not None and r["ee_pose"].shape == (7,) and abs(np.linalg.norm(r["ee_pose"][3:]) - 1) < 1e-6
    assert r["key_points_pose"] is not None and r["icp_stats"][2] >= 1


def test_heads_vs_reference_golden(golden):
    """oracle restatements of utils/output.py:45-87 against outputs of the reference's own functions
    (tests/golden/make_golden.py imports utils/output.py with the oracle package standing in for MinkowskiEngine)."""
    for th, tag in ((0.75, "75"), (0.999, "999")):
        idx, cls, pr = G.key_point_predictions(golden["kp_logits"], th)
        assert np.array_equal(cls, golden["kp_cls" + tag]) and np.array_equal(idx, golden["kp_idx" + tag])
        assert np.allclose(np.asarray(pr), golden["kp_pr" + tag], atol=1e-7)
    idx, cls, pr = G.key_point_predictions(golden["kp10_logits"], 0.75)
    assert np.array_equal(cls, golden["kp10_cls"]) and np.array_equal(idx, golden["kp10_idx"])
    assert np.allclose(G.pred_center(golden["vote_out"], golden["vote_coords"], ee_r=0.02), golden["vote_center"],
                       atol=1e-6)
    assert np.allclose(G.pred_center(golden["vote_out"], golden["vote_coords"], ee_r=0.02, q=golden["vote_q"]),
                       golden["vote_center_q"], atol=1e-6)
    assert np.array_equal(G.segmentation_labels(golden["seg_logits"]), golden["seg_preds"])


def test_pose_metric_vs_reference_golden(golden):
    """the tolerance metric of the pose parity tests (rotation angle, translation distance) against
    utils/metrics.py:110-127 compute_pose_metrics."""
    for i, (p, p2) in enumerate(zip(golden["pose"], golden["pose2"])):
        Ra = G.quaternion_rotation_matrix(p[3:], switch_w=False)
        Rb = G.quaternion_rotation_matrix(p2[3:], switch_w=False)
        assert abs(np.radians(G.rotation_angle_deg(Ra, Rb)) - golden["metric_angle"][i]) < 1e-6, i
        assert abs(np.linalg.norm(p[:3] - p2[:3]) - golden["metric_dist"][i]) < 1e-12


def test_translation_magic_vs_reference_golden(golden):
    """a17b against InferenceEngine.predict_translation of the reference itself (app/inference_engine.py:459-489)."""
    for i, n in enumerate(golden["trans_n"]):
        got = G.translation_magic(golden["trans_pts"][i][:n], golden["trans_q"][i])
        assert np.allclose(got, golden["trans_out"][i], atol=1e-6), (i, got, golden["trans_out"][i])


def test_roi_mask_vs_reference_golden(golden):
    lim = golden["roi_limits"]
    assert np.array_equal(G.roi_mask(golden["roi_pts"], *lim), golden["roi_mask"])
    assert np.array_equal(G.roi_mask(golden["roi_pts"]), golden["roi_mask_default"])
    assert np.array_equal(G.roi_mask(golden["roi_pts"], *lim, offset=0.1), golden["roi_mask_offset"])
    assert 0.2 < golden["roi_mask"].mean() < 0.8


def test_calibrate_vs_reference_golden(golden, built_lib):
    """the calibration tail (b200calib.calibration.calibrate) against InferenceEngine.calibrate of the reference
    itself (app/inference_engine.py:152-244): several positions, unconfident frames, frames without key points."""
    from b200calib import calibration as C
    from b200calib.pipeline import FrameResult
    rows = golden["calib_rows"]
    data = {}
    for r in rows:
        kp = None if np.isnan(r[16]) else r[16:23]
        kpb = None if np.isnan(r[23]) else r[23:30]
        fr = FrameResult(segmentation=None, ee_pose=r[2:9], base_pose=r[9:16], key_points_pose=kp,
                         key_points_base_pose=kpb, is_confident=bool(r[1]))
        data.setdefault(str(int(r[0])), []).append(fr)
    cl = golden["calib_camera_link"]

    def close(a, b):
        a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
        return np.allclose(a[:3], b[:3], atol=1e-9) and min(np.abs(a[3:] - b[3:]).max(), np.abs(a[3:] + b[3:]).max()) < 1e-9
    allr = C.calibrate(data, camera_link_transformation_pose=cl)
    assert close(allr.pose_camera_link, golden["calib_pose"]) and close(allr.base_pose, golden["calib_base"])
    assert close(allr.key_points_base_pose, golden["calib_kp_base"])
    assert close(allr.base_pose_camera_link, golden["calib_base_cl"])
    one = C.calibrate({"0": data["0"]}, camera_link_transformation_pose=cl)
    assert close(one.pose_camera_link, golden["calib_one_pose"])
    assert close(one.base_pose_camera_link, golden["calib_one_base_cl"])
    assert C.calibrate({"0": data["0"][:1]}) is None   # fewer than 2 confident frames
