"""Pose averaging with the names of utils/calibration.py (:15-139). O(frames) host math in float64:
the tail of the pipeline that runs on rank 0 after the all-gather."""
import numpy as np


def get_outliers(y, m=2.0):
    y = np.asarray(y)
    d = np.abs(y - np.median(y))
    mdev = np.median(d)
    s = d / mdev if mdev else np.zeros_like(d)
    is_outlier = s > m
    return is_outlier, y[is_outlier]


def remove_pose_outliers(poses):
    # the reference computes the outlier flags and then returns the input unchanged (utils/calibration.py:55-61)
    return poses


def compute_quaternions_weighted_average(Q, w):
    """Markley eigen-average: principal eigenvector of sum_i w_i q_i q_i^T / sum w (utils/calibration.py:69-95)."""
    Q = np.asarray(Q)   # kept in the caller's dtype: the reference forms q q^T in float32 when the poses are float32
    w = np.asarray(w, dtype=np.float64)
    A = np.zeros((4, 4))
    for i in range(Q.shape[0]):
        A = w[i] * np.outer(Q[i], Q[i]) + A
    A = (1.0 / np.sum(w)) * A
    vals, vecs = np.linalg.eig(A)
    vecs = vecs[:, vals.argsort()[::-1]]
    return np.real(vecs[:, 0])


def compute_quaternions_average(Q):
    return compute_quaternions_weighted_average(Q, np.ones(len(Q)))


def compute_translations_average(t, weights=None):
    t = np.asarray(t)
    if weights is None:
        weights = np.ones(len(t))
    return np.sum(t * weights.reshape(-1, 1), axis=0) / np.sum(weights)


def compute_poses_average(poses, weights=None):
    """poses [N,7] x,y,z,qw,qx,qy,qz (utils/calibration.py:117-139)."""
    if poses is None or len(poses) == 0:
        return poses
    poses = np.asarray(poses)
    if poses.ndim != 2:
        poses = np.array(poses.reshape(-1, 7), copy=True)
    if len(poses) == 1:
        return poses[0]
    if weights is None or len(weights) != len(poses):
        weights = np.ones(len(poses))
    out = np.zeros(7)
    out[:3] = compute_translations_average(poses[:, :3], weights=weights)
    out[3:] = compute_quaternions_weighted_average(poses[:, 3:], weights)
    return out


# ---- InferenceEngine.calibrate / _calibrate_individual (app/inference_engine.py:152-244) -------------------------
class CalibrationResult:
    """CalibrationResultDTO / TestResultDTO fields (app/dto.py:50-70) that the calibration tail fills."""

    def __init__(self):
        self.is_confident = True
        self.ee_pose = self.base_pose = self.key_points_pose = self.key_points_base_pose = None
        self.base_pose_camera_link = self.key_points_base_pose_camera_link = None
        self.pose_camera_link = None


def _stack32(rows):
    return np.array(rows, dtype=np.float32)


def calibrate_individual(data, weights=None, confident_count=2, camera_link_transformation_pose=None):
    """average of the confident results of ONE robot position (frames) or of several positions (their averages).
    `data`: objects with is_confident, ee_pose, base_pose, key_points_pose, key_points_base_pose (FrameResult /
    CalibrationResult). Poses are stacked as float32 and averaged in float64 like the reference."""
    from .transformation import transform_pose2pose
    out = CalibrationResult()
    try:
        conf = [d for d in data if d.is_confident]
        if len(conf) < confident_count:
            return None
        if weights is not None:
            weights = np.asarray(weights)[np.array([d.is_confident for d in data], dtype=bool)]
        out.ee_pose = compute_poses_average(remove_pose_outliers(_stack32([d.ee_pose for d in conf])), weights=weights)
        out.base_pose = compute_poses_average(remove_pose_outliers(_stack32([d.base_pose for d in conf])),
                                              weights=weights)
        out.key_points_pose = compute_poses_average(
            remove_pose_outliers(_stack32([d.key_points_pose for d in conf if d.key_points_pose is not None])),
            weights=weights)
        out.key_points_base_pose = compute_poses_average(
            remove_pose_outliers(_stack32([d.key_points_base_pose for d in conf if d.key_points_base_pose is not None])),
            weights=weights)
        bl = kl = None
        # reference quirk (:225-232): TestResultDTO subclasses ResultDTO, so the `isinstance(..., ResultDTO)` branch is
        # taken at BOTH levels: the camera-link poses are always recomputed from (averaged) base poses, the branch that
        # would average the per-position camera-link poses is dead code.
        if camera_link_transformation_pose is not None:
            cl = np.asarray(camera_link_transformation_pose, dtype=np.float32)
            bl = _stack32([transform_pose2pose(d.base_pose, cl) for d in conf if d.base_pose is not None])
            kl = _stack32([transform_pose2pose(d.key_points_base_pose, cl) for d in conf
                           if d.key_points_base_pose is not None])
        if bl is not None:
            out.base_pose_camera_link = compute_poses_average(remove_pose_outliers(bl), weights=weights)
        if kl is not None:
            out.key_points_base_pose_camera_link = compute_poses_average(remove_pose_outliers(kl), weights=weights)
    except Exception:  # the reference swallows every failure and marks the result unconfident (:242-243)
        out.is_confident = False
    return out


def calibrate(data, camera_link_transformation_pose=None):
    """InferenceEngine.calibrate (:152-194). data: dict position -> list of per-frame results. Returns a
    CalibrationResult whose pose_camera_link is the average of {base_pose, key_points_base_pose}, or None."""
    individual = [calibrate_individual(v, camera_link_transformation_pose=camera_link_transformation_pose)
                  for v in data.values()]
    individual = [v for v in individual if v is not None]
    if len(data) == 1 and len(individual) > 0:
        raw = individual[0]
    else:
        raw = calibrate_individual(individual, camera_link_transformation_pose=camera_link_transformation_pose)
        if raw is None:
            return None
    raw.pose_camera_link = compute_poses_average(np.stack((raw.base_pose, raw.key_points_base_pose), axis=0))
    return raw
