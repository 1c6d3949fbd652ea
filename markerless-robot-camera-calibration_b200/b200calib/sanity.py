"""Host-side mirror of the reference's per-frame sanity check:

  InferenceEngine.check_sanity          app/inference_engine.py:246-279
  get_6_key_points ("gt" key points)    utils/data.py:255-335 (+ get_closest_point :125-138, center_at_origin
                                        utils/preprocess.py:8-11)
  compute_kp_error                      utils/metrics.py:130-136

A few thousand EE points per frame, a handful of reductions: NumPy on the host (the reference does the same after its
device->host copies); the pipeline already has the EE crop and the poses on the host when the check runs.
Same names, argument meaning and return values as the reference functions; float64 arithmetic like NumPy's defaults
there. Pinned by outputs of the reference's own functions (tests/golden/make_golden_sanity.py).
"""
import numpy as np

from .transformation import get_quaternion_rotation_matrix

# EE-frame templates of utils/data.py:264-271 (corner / gripper key points) and :280-285 (far "bounding box" probes
# whose nearest EE points are the corner candidates)
KEY_POINT_TEMPLATE = np.array([[0.02, 0.09, 0.0], [0.01, -0.1, 0.0], [0.014, 0.095, 0.07], [0.014, -0.095, 0.07],
                               [0.0, 0.048, 0.12], [0.0, -0.048, 0.12]], dtype=np.float64)
CORNER_PROBES = np.array([[0.24, 0.32, -0.2], [0.24, -0.32, -0.2], [0.24, 0.32, 0.2], [0.24, -0.32, 0.2]],
                         dtype=np.float64)


def get_closest_point(p, points, maximize_dim=None):
    """utils/data.py:125-138: nearest of `points` to p; with maximize_dim the probe is first moved to the largest
    coordinate of the points along that axis. Returns (index, point, distance) or None for an empty set."""
    if len(points) < 1:
        return None
    probe = np.array(p, dtype=np.float64)
    if maximize_dim is not None:
        probe[maximize_dim] = points[:, maximize_dim].max()
    d = np.linalg.norm(points - probe, axis=1)
    i = int(d.argmin())
    return i, points[i], d[i]


def get_6_key_points(ee_points, pose, switch_w=True, euclidean_threshold=0.03, ignore_label=-100):
    """utils/data.py:255-335. EE points and the pose position go to the EE frame (R^T p), centred on the pose
    position; the four corner key points are the EE points nearest to four far probes (accepted when within
    `euclidean_threshold` of the template corner), the two gripper key points are the points nearest to a probe lifted
    to the largest z of each gripper side (mirrored when one side is empty, both set to the larger z). Returns the
    key points back in the camera frame [6,3] and the index of the EE point each one sits on (ignore_label: template)."""
    ee_points = np.asarray(ee_points)
    pose = np.asarray(pose)
    R = get_quaternion_rotation_matrix(pose[3:], switch_w=switch_w)
    local = np.concatenate((ee_points, pose[:3].reshape(1, 3))) @ R  # rows = (R^T p)^T
    origin = local[-1].copy()  # centre_at_origin of the single pose point = the point itself
    local = local[:-1] - origin

    kps = KEY_POINT_TEMPLATE.copy()
    point_idx = np.full(len(kps), ignore_label, dtype=np.int64)

    front = (local[:, 0] > -0.005) & (local[:, 2] < 0.09)
    front_idx = np.flatnonzero(front)
    if len(front_idx) < 1:
        return np.array([]), np.array([])
    sel = local[front_idx]
    nearest = np.linalg.norm(CORNER_PROBES[:, None, :] - sel[None, :, :], axis=2).argmin(axis=1)
    cand_idx = front_idx[nearest]
    cand = local[cand_idx]
    ok = np.linalg.norm(kps[:4] - cand, axis=1) < euclidean_threshold
    kps[:4][ok] = cand[ok]
    point_idx[:4][ok] = cand_idx[ok]

    grip_idx = np.flatnonzero(local[:, 2] > 0.08)
    grip = local[grip_idx]
    found = [None, None]
    for side, (probe, keep) in enumerate((([0, 0.01, 0.1], grip[:, 1] > 0), ([0, -0.01, 0.1], grip[:, 1] < 0))):
        if keep.any():
            i, pt, _ = get_closest_point(probe, grip[keep], maximize_dim=2)
            found[side] = pt
            kps[4 + side] = pt
            # the reference indexes the list of ALL gripper points with the index found inside the one-sided subset
            # (utils/data.py:309,321); kept as is — only the first four entries are consumed (check_sanity)
            point_idx[4 + side] = grip_idx[i]
    if found[0] is None and found[1] is not None:
        kps[4] = found[1] * [1, -1, 1]
    elif found[0] is not None and found[1] is None:
        kps[5] = found[0] * [1, -1, 1]
    kps[4, 2] = kps[5, 2] = max(kps[4, 2], kps[5, 2])

    kps = (kps + origin) @ R.T
    return kps, point_idx


def compute_kp_error(gt_coords, kp_coords, kp_classes):
    """utils/metrics.py:130-136: mean distance between the predicted key points and the gt key points of their classes;
    100 when fewer than two are given."""
    if len(gt_coords) < 2 or len(kp_coords) < 2 or len(kp_classes) < 2:
        return 100
    return np.linalg.norm(gt_coords[kp_classes] - kp_coords, axis=1).mean()


def check_sanity(points, segmentation, ee_pose, key_points, min_num_of_ee_points=2048, kp_error_margin=0.05):
    """app/inference_engine.py:246-279 with the thresholds of config/default.yaml:148-149,185. points [N,3] and the
    per-point labels of the frame (2 = EE), the predicted EE pose (x,y,z,qw,qx,qy,qz), the predicted key points as
    (class, xyz) pairs (ResultDTO.key_points). False when the EE has too few points, when a corner of the EE cannot be
    found where the pose says it should be, or when the predicted key points are further than the margin from them."""
    segmentation = np.asarray(segmentation)
    ee_mask = segmentation == 2
    if ee_mask.sum() < min_num_of_ee_points:
        return False
    ee_raw_points = np.asarray(points)[ee_mask]
    kp_gt_coords, kp_gt_classes = get_6_key_points(ee_raw_points, np.asarray(ee_pose), switch_w=False,
                                                   euclidean_threshold=0.04)
    if any(kp_gt_classes[:4] < 0):
        return False
    if len(key_points) > 3:
        kp_pred_classes = np.array([c for c, _ in key_points], dtype=np.int64)
        kp_pred_coords = np.array([xyz for _, xyz in key_points], dtype=np.float32)
        if compute_kp_error(kp_gt_coords, kp_pred_coords, kp_pred_classes) > kp_error_margin:
            return False
    return True
