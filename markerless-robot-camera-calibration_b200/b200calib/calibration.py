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
    Q = np.asarray(Q, dtype=np.float64)
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
