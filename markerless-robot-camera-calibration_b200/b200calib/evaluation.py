"""Evaluation record of a run: what the reference's app/test.py:78-290 computes per frame, per robot
position and overall, written as JSON instead of its xlsx sheet (:292-509).

Per frame (instance): segmentation accuracy / precision / recall (utils/metrics.py:50-107), translation and rotation
error of the network pose ("nn"), of its ICP refinement ("nn_icp"), of the key-point pose ("kp", "kp_icp")
(utils/metrics.py:110-127), ADD of each (utils/metrics.py:139-151), mean key-point error (utils/metrics.py:130-136),
base->camera error against the known camera pose; then the mean per position, the mean of the position means
(statistics.mean, like the reference) and the error of the final calibration (app/test.py:279-286).
Metric functions keep the reference's names and arithmetic and are pinned by outputs of the reference's own functions
(tests/golden/make_golden_eval.py)."""
import json
import statistics
from collections import defaultdict

import numpy as np

from .sanity import compute_kp_error, get_6_key_points
from .transformation import get_quaternion_rotation_matrix

CLASSES = ("background", "arm", "ee")   # config/default.yaml INFERENCE.SEGMENTATION.classes


def _qmul(q, r):
    """utils/quaternion.py qmul_np for single w,x,y,z quaternions (Hamilton product)."""
    w = q[0] * r[0] - q[1] * r[1] - q[2] * r[2] - q[3] * r[3]
    x = q[0] * r[1] + q[1] * r[0] + q[2] * r[3] - q[3] * r[2]
    y = q[0] * r[2] - q[1] * r[3] + q[2] * r[0] + q[3] * r[1]
    z = q[0] * r[3] + q[1] * r[2] - q[2] * r[1] + q[3] * r[0]
    return np.array([w, x, y, z])


def compute_pose_metrics(gt, pred):
    """utils/metrics.py:110-127; poses x,y,z,qw,qx,qy,qz -> dict(dist_position [m], angle_diff [rad])."""
    gt, pred = np.asarray(gt, dtype=np.float64), np.asarray(pred, dtype=np.float64)
    gt_rot = gt[3:] / np.linalg.norm(gt[3:])
    pred_rot = pred[3:] / np.linalg.norm(pred[3:])
    q_mul = _qmul(gt_rot, pred_rot * np.array([1.0, -1.0, -1.0, -1.0]))
    angle = np.abs(2 * np.arctan2(np.linalg.norm(q_mul[1:]), q_mul[0]))
    return dict(dist_position=float(np.linalg.norm(gt[:3] - pred[:3])), angle_diff=float(min(angle, 2 * np.pi - angle)))


def compute_segmentation_metrics(gt, pred, classes=CLASSES):
    """utils/metrics.py:50-107: per class accuracy / precision / recall (precision, recall = 1 when there is no false
    positive / negative), overall accuracy = (sensitivity + specificity) / 2 over the summed confusion counts, overall
    precision / recall = mean over classes."""
    gt, pred = np.asarray(gt), np.asarray(pred)
    res = dict(class_results={})
    precisions, recalls = [], []
    tp_s = tn_s = fp_s = fn_s = 0
    for ci, cn in enumerate(classes):
        g, p = gt == ci, pred == ci
        tp = int((g & p).sum())
        tn = int(len(gt) - (g | p).sum())
        fp = int(p.sum()) - tp
        fn = int(g.sum()) - tp
        tp_s, tn_s, fp_s, fn_s = tp_s + tp, tn_s + tn, fp_s + fp, fn_s + fn
        precision = int(fp == 0) or tp / (tp + fp)
        recall = int(fn == 0) or tp / (tp + fn)
        res["class_results"][cn] = dict(accuracy=(tp + tn) / (tp + tn + fp + fn), precision=precision, recall=recall)
        precisions.append(precision)
        recalls.append(recall)
    res["accuracy"] = (tp_s / (tp_s + fn_s) + tn_s / (tn_s + fp_s)) / 2
    res["precision"] = statistics.mean(precisions)
    res["recall"] = statistics.mean(recalls)
    return res


def compute_ADD_np(points, gt_pose, pred_pose):
    """utils/metrics.py:139-151: mean distance between the model points under the two poses."""
    points = np.asarray(points)
    gt_pose, pred_pose = np.asarray(gt_pose), np.asarray(pred_pose)
    Rg = get_quaternion_rotation_matrix(gt_pose[3:], switch_w=False)
    Rp = get_quaternion_rotation_matrix(pred_pose[3:], switch_w=False)
    a = (Rg @ points.T) + gt_pose[:3].reshape(3, 1)
    b = (Rp @ points.T) + pred_pose[:3].reshape(3, 1)
    return float(np.linalg.norm(a - b, axis=0).mean())


def evaluate_frame(points, gt_labels, gt_pose, result, stages, gt_base2cam=None, evaluate_segmentation=True):
    """one instance record of app/test.py:78-229. gt_pose: x,y,z,qw,qx,qy,qz; `result`: FrameResult; `stages`: dict
    with the poses before the ICP refinement (ee_pose_initial, kp_pose_initial), when the pipeline returned them."""
    rec = {}
    points = np.asarray(points)
    gt_labels = np.asarray(gt_labels)
    if evaluate_segmentation:
        rec["segmentation"] = compute_segmentation_metrics(gt_labels, result.segmentation)
    if result.ee_pose is None:
        return rec
    ee_gt = points[gt_labels == 2]
    if len(ee_gt) < 1:
        ee_gt = points[[1, 2, 3]]                                   # app/test.py:120-121
    centered = ee_gt - (ee_gt.max(axis=0) + ee_gt.min(axis=0)) / 2  # centre_at_origin of the GT EE points (:123-124)
    rec["dist_position"], rec["angle_diff"] = {}, {}
    named = [("nn", stages.get("ee_pose_initial")), ("nn_icp", result.ee_pose),
             ("kp", stages.get("kp_pose_initial")), ("kp_icp", result.key_points_pose)]
    for name, pose in named:
        if pose is None:
            continue
        m = compute_pose_metrics(gt_pose, pose)
        rec["dist_position"][name] = m["dist_position"]
        rec["angle_diff"][name] = m["angle_diff"]
        rec["ADD_" + name] = compute_ADD_np(centered, gt_pose, pose)
    ee_pred = points[np.asarray(result.segmentation) == 2] if stages.get("ee_points") is None else stages["ee_points"]
    if len(ee_pred) and result.key_points:
        kp_gt, _ = get_6_key_points(ee_pred, np.asarray(gt_pose), switch_w=False)
        cls = np.array([c for c, _ in result.key_points], dtype=np.int64)
        xyz = np.array([p for _, p in result.key_points], dtype=np.float32)
        rec["mean_kp_error"] = float(compute_kp_error(kp_gt, xyz, cls))
    if gt_base2cam is not None:
        rec["base2cam"] = {}
        if result.base_pose is not None:
            m = compute_pose_metrics(gt_base2cam, result.base_pose)
            rec["base2cam"].update(dist_position=m["dist_position"], angle_diff=m["angle_diff"])
        if result.key_points_base_pose is not None:
            m = compute_pose_metrics(gt_base2cam, result.key_points_base_pose)
            rec["base2cam"].update(dist_position_kp=m["dist_position"], angle_diff_kp=m["angle_diff"])
    rec["is_confident"] = bool(result.is_confident)
    return rec


def aggregate(instances, calibration_pose=None, gt_base2cam=None):
    """app/test.py:239-286: instance records (dicts with a 'position' key) -> per-position lists -> the mean per
    position -> the mean of the position means; plus the error of the final calibration pose."""
    by_pos = defaultdict(list)
    for rec in instances:
        by_pos[rec.get("position", "p0")].append(rec)
    position_results = {}
    for pos, recs in by_pos.items():
        pr = defaultdict(list)
        for r in recs:
            for name, v in r.get("dist_position", {}).items():
                pr["dist_position_" + name].append(v)
            for name, v in r.get("angle_diff", {}).items():
                pr["angle_diff_" + name].append(v)
            for k, v in r.items():
                if k.startswith("ADD_") or k == "mean_kp_error":
                    pr[k].append(v)
            for k, v in r.get("base2cam", {}).items():
                pr["base2cam_" + k].append(v)
            seg = r.get("segmentation")
            if seg is not None:
                for k in ("accuracy", "precision", "recall"):
                    pr["segmentation_" + k].append(seg[k])
                    for cn, cr in seg["class_results"].items():
                        pr[f"segmentation_{cn}_{k}"].append(cr[k])
        position_results[pos] = {k: [float(x) for x in v] for k, v in pr.items()}
    overall = defaultdict(list)
    for pr in position_results.values():
        for k, v in pr.items():
            if len(v) > 0:
                overall[k].append(statistics.mean(v))
    overall = {k: float(statistics.mean(v)) for k, v in overall.items()}
    overall["calibration_angle_diff"] = overall["calibration_dist_position"] = -100
    if calibration_pose is not None and gt_base2cam is not None:
        m = compute_pose_metrics(calibration_pose, gt_base2cam)
        overall["calibration_angle_diff"], overall["calibration_dist_position"] = m["angle_diff"], m["dist_position"]
    return dict(positions=position_results, overall=overall, frames=len(instances),
                frames_confident=int(sum(1 for r in instances if r.get("is_confident"))))


def write_report(path, report):
    """the JSON counterpart of export_to_xslx (app/test.py:292-509): millimetres / degrees are left to the reader, the
    record keeps SI units (metres, radians) like the reference's in-memory results."""
    with open(path, "w") as fp:
        json.dump(report, fp, indent=1, sort_keys=True)
    return path
