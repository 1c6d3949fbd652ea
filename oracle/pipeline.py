"""Oracle (TEST INFRASTRUCTURE) of the per-frame inference pipeline: a CPU restatement of
app/inference_engine.py InferenceEngine.predict (:281-382) with predict_segmentation (:395-435),
predict_rotation (:437-457), predict_translation (:459-489), predict_key_points ME branch (:539-555),
predict_pose_from_kp (:384-393) and match_icp (utils/icp.py:50-81), one frame at a time, fp32 networks on the
oracle MinkowskiEngine package, float64 geometry — the way the reference runs it.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
Parity is UNPINNED for the rows whose arithmetic lives in absent third-party code.
"""
import numpy as np
import torch

from . import MinkowskiEngine as OME
from . import geometry as og

# app/inference_engine.py:128-137
REFERENCE_KEY_POINTS = np.array([
    [0.01982731, 0.08085986, 0.00321919],
    [0.02171595, -0.08986182, 0.00388430],
    [0.01288678, 0.09103118, 0.06127814],
    [0.02079032, -0.09790908, 0.05609143],
    [-0.00185802, 0.04654205, 0.11564558],
    [0.00241113, -0.04262756, 0.11564558],
])

DEFAULTS = dict(seg_scale=200.0, rot_scale=200.0, kp_scale=800.0, ee_point_counts_threshold=512, cluster_dist=0.06,
                kp_conf_threshold=0.75, icp_enabled=True, rot_center_at_origin=True, kp_center_at_origin=True,
                translation_x_offset=-0.015)


def normalize_colors(rgb_input):
    """utils/preprocess.py:20-37: /255 when the maximum exceeds 2; per-channel min-max rescale to [0, 1] when a value
    is negative (sklearn.preprocessing.minmax_scale: X * scale + min_ with scale = 1 / (max - min) in the input's
    float32); - 0.5 when the result lies in [0, 1]."""
    rgb = np.array(rgb_input, copy=True)
    if not rgb.size:
        return rgb
    if rgb.max() > 2:
        rgb /= 255.0
    if rgb.min() < 0:
        for c in range(3):
            col = rgb[:, c]
            lo, hi = col.min(), col.max()
            rng = hi - lo
            scale = (np.float32(1.0) / (rng if rng != 0 else np.float32(1.0))).astype(rgb.dtype)
            min_ = (0 - lo * scale).astype(rgb.dtype)
            col *= scale
            col += min_
    if rgb.min() > (-1e-6) and rgb.max() < (1 + 1e-6):
        rgb -= 0.5
    return rgb


def center_at_origin(points):
    """utils/preprocess.py:8-11."""
    off = (points.max(axis=0) + points.min(axis=0)) / 2
    return points - off, off


def _field(points_f32, feats, scale):
    pts = torch.from_numpy(np.ascontiguousarray(points_f32)).to(torch.float32)
    return OME.TensorField(features=feats, coordinates=OME.utils.batched_coordinates([pts * scale], dtype=torch.float32),
                           quantization_mode=OME.SparseTensorQuantizationMode.UNWEIGHTED_AVERAGE,
                           minkowski_algorithm=OME.MinkowskiAlgorithm.SPEED_OPTIMIZED)


def predict_segmentation(seg_model, points, rgb_norm, scale, cluster_dist=0.06):
    """app/inference_engine.py:395-435 -> (labels [N] int64 after the EE cluster filter, dict with the raw arg-max
    labels, the per-point top-2 logit margin and the logit scale, which parity tests use to skip undecided points)."""
    fld = _field(points, torch.from_numpy(rgb_norm).to(torch.float32), scale)
    out = seg_model(fld.sparse()).slice(fld)
    seg = og.segmentation_labels(out.F).astype(np.int64)
    top2 = out.F.float().topk(2, dim=1)[0]
    raw = dict(labels=seg.copy(), margin=(top2[:, 0] - top2[:, 1]).numpy(), scale=float(out.F.abs().max()),
               point_logits=out.F.float().numpy())
    ee_idx = np.where(seg == 2)[0]
    seg[ee_idx] = 1
    if len(ee_idx) > 1:
        inside = og.largest_cluster(points[ee_idx].astype(np.float32), cluster_dist)
        seg[ee_idx[inside]] = 2
    return seg, raw


def filter_ee(points, labels, cluster_dist=0.06):
    """the relabel + largest-cluster step of predict_segmentation on given arg-max labels."""
    seg = np.array(labels, dtype=np.int64, copy=True)
    ee_idx = np.where(seg == 2)[0]
    seg[ee_idx] = 1
    if len(ee_idx) > 1:
        inside = og.largest_cluster(points[ee_idx].astype(np.float32), cluster_dist)
        seg[ee_idx[inside]] = 2
    return seg


def predict_frame(models, cad, points, rgb, cfg=None, gt_labels=None, ee2base_pose=None):
    """-> dict(segmentation, ee_pose, key_points_pose, key_points, icp_stats, ...) for one frame.
    models: dict(seg=..., rot=..., kp=...) of oracle-ME networks (rot / kp may be None).
    gt_labels: use these arg-max labels for the EE crop instead of the network's (random-init weights)."""
    c = dict(DEFAULTS)
    c.update(cfg or {})
    res = dict(ee_pose=None, key_points_pose=None, key_points=None, icp_stats=None, kp_icp_stats=None,
               base_pose=None, key_points_base_pose=None)
    points = np.asarray(points, dtype=np.float32)
    with torch.no_grad():
        rgbn = normalize_colors(np.asarray(rgb, dtype=np.float32))
        seg, raw = predict_segmentation(models["seg"], points, rgbn, c["seg_scale"], c["cluster_dist"])
        res["segmentation_pred"], res["segmentation_raw"] = seg, raw
        if gt_labels is not None:
            seg = filter_ee(points, gt_labels, c["cluster_dist"])
        res["segmentation"] = seg
        ee_idx = np.where(seg == 2)[0]
        if len(ee_idx) < c["ee_point_counts_threshold"] or models.get("rot") is None:
            return res
        ee_raw = points[ee_idx]
        ee_rgb = torch.from_numpy(rgbn[ee_idx]).to(torch.float32)
        # rotation (:437-457)
        rot_pts = center_at_origin(ee_raw)[0] if c["rot_center_at_origin"] else ee_raw
        q = models["rot"](_field(rot_pts, ee_rgb, c["rot_scale"]).sparse())[0][3:7].numpy()
        # translation (:459-489)
        pos = og.translation_magic(ee_raw, q, c["translation_x_offset"])
        ee_pose = np.concatenate((pos, q.astype(np.float64)))
        # key points (:539-555) + Kabsch (:384-393)
        kp_pose = None
        if models.get("kp") is not None:
            kp_pts = center_at_origin(ee_raw)[0] if c["kp_center_at_origin"] else ee_raw
            kfld = _field(kp_pts, ee_rgb, c["kp_scale"])
            kout = models["kp"](kfld.sparse()).slice(kfld).F
            kp_idx, kp_classes, kp_probs = og.key_point_predictions(kout, c["kp_conf_threshold"])
            res["key_points"] = (kp_idx, kp_classes, np.asarray(kp_probs))
            if len(kp_classes) >= 4:
                R, t = og.rigid_transform_3D(REFERENCE_KEY_POINTS[kp_classes], ee_raw[kp_idx])
                kp_pose = np.concatenate((t, og.q_from_matrix(R)))
        # ICP (:358-362)
        if c["icp_enabled"] and cad is not None:
            T, fit, rmse, it = og.icp_point_to_point(cad, ee_raw, og.transformation_matrix(ee_pose))
            ee_pose = og.pose_from_matrix(T)
            res["icp_stats"] = np.array([fit, rmse, it])
            res["ee_T"] = T
            if kp_pose is not None:
                Tk, fk, rk, ik = og.icp_point_to_point(cad, ee_raw, og.transformation_matrix(kp_pose))
                kp_pose = og.pose_from_matrix(Tk)
                res["kp_icp_stats"] = np.array([fk, rk, ik])
                res["kp_T"] = Tk
        res["ee_pose"], res["key_points_pose"] = ee_pose, kp_pose
        if ee2base_pose is not None:
            inv = np.linalg.inv(og.transformation_matrix(ee2base_pose))
            res["base_pose"] = og.pose_from_matrix(og.transformation_matrix(ee_pose) @ inv)
            if kp_pose is not None:
                res["key_points_base_pose"] = og.pose_from_matrix(og.transformation_matrix(kp_pose) @ inv)
    return res
