"""Pose algebra with the names and argument meaning of utils/transformation.py.

get_rigid_transform_3D (utils/transformation.py:178-222) runs on the GPU (K9, batched); the remaining
functions are O(1) host math on 4x4 matrices / quaternions, kept in NumPy (float64) like the reference.
Quaternions are W,X,Y,Z in memory."""
import ctypes

import numpy as np
import torch
from scipy.spatial.transform import Rotation

from MinkowskiEngine._lib import lib, check, ptr, stream
from MinkowskiEngine.core import _count


def switch_w(pose):
    """x,y,z,qx,qy,qz,qw -> x,y,z,qw,qx,qy,qz (utils/transformation.py:7-13)."""
    pose = np.asarray(pose)
    return np.concatenate((pose[:-4], pose[-1:], pose[-4:-1]))


def get_quaternion_rotation_matrix(Q_init, switch_w=True):
    """utils/transformation.py:16-60; arithmetic stays in the dtype of the input like the reference."""
    Q = np.asarray(Q_init)
    if switch_w:
        Q = np.concatenate((Q[-1:], Q[:3]))
    q0, q1, q2, q3 = Q[0], Q[1], Q[2], Q[3]
    return np.array([[2 * (q0 * q0 + q1 * q1) - 1, 2 * (q1 * q2 - q0 * q3), 2 * (q1 * q3 + q0 * q2)],
                     [2 * (q1 * q2 + q0 * q3), 2 * (q0 * q0 + q2 * q2) - 1, 2 * (q2 * q3 - q0 * q1)],
                     [2 * (q1 * q3 - q0 * q2), 2 * (q2 * q3 + q0 * q1), 2 * (q0 * q0 + q3 * q3) - 1]])


def get_transformation_matrix(pose, switch_w=False):
    pose = np.asarray(pose)
    T = np.eye(4)
    T[:3, :3] = get_quaternion_rotation_matrix(pose[3:], switch_w=switch_w)
    T[:3, 3] = pose[:3]
    return T


def get_transformation_matrix_inverse(trans_mat):
    T = np.array(trans_mat, copy=True)
    T[:3, :3] = trans_mat[:3, :3].T
    T[:3, 3] = -T[:3, :3] @ trans_mat[:3, 3]
    return T


def get_q_from_matrix(rot_mat):
    q = Rotation.from_matrix(np.array(rot_mat, copy=True)).as_quat()  # x,y,z,w; sign not canonicalised
    return np.concatenate((q[3:], q[:3]))


def get_pose_from_matrix(trans_mat):
    return np.concatenate((trans_mat[:3, 3], get_q_from_matrix(trans_mat[:3, :3])))


def get_poses_from_matrices(trans_mats):
    """batched get_pose_from_matrix: [S,4,4] -> [S,7] (x,y,z,qw,qx,qy,qz), one SciPy call for all frames."""
    T = np.asarray(trans_mats, dtype=np.float64)
    if len(T) == 0:
        return np.zeros((0, 7))
    q = Rotation.from_matrix(T[:, :3, :3].copy()).as_quat()  # x,y,z,w
    return np.concatenate((T[:, :3, 3], q[:, 3:], q[:, :3]), axis=1)


def get_pose_inverse(pose):
    return get_pose_from_matrix(get_transformation_matrix_inverse(get_transformation_matrix(pose)))


def get_base2cam_matrix(ee2cam_pose, ee2robot_pose):
    return get_transformation_matrix(ee2cam_pose) @ get_transformation_matrix_inverse(
        get_transformation_matrix(ee2robot_pose))


def get_base2cam_pose(ee2cam_pose, ee2robot_pose):
    return get_pose_from_matrix(get_base2cam_matrix(ee2cam_pose, ee2robot_pose))


def transform_pose2pose_matrix(pose1, pose2):
    return get_transformation_matrix(pose1) @ get_transformation_matrix(pose2)


def transform_pose2pose(pose1, pose2):
    return get_pose_from_matrix(transform_pose2pose_matrix(pose1, pose2))


# ------------------------------------------------------------------------------------------------ K9
def rigid_transform_3D_batched(reference, target, npairs=None):
    """Batched Kabsch on the GPU. reference/target: [P, kmax, 3] float64 CUDA tensors (rows beyond npairs[p]
    ignored). Returns R [P,3,3], t [P,3] float64 CUDA tensors with R @ reference_i + t ~= target_i."""
    reference = reference.to(torch.float64).contiguous()
    target = target.to(torch.float64).contiguous()
    P, kmax, _ = reference.shape
    dev = reference.device
    if npairs is None:
        npairs = torch.full((P,), kmax, dtype=torch.int32, device=dev)
    npairs = npairs.to(torch.int32).contiguous()
    R = torch.empty((P, 9), dtype=torch.float64, device=dev)
    t = torch.empty((P, 3), dtype=torch.float64, device=dev)
    check(lib.b2me_kabsch_batched(ptr(reference), ptr(target), ptr(npairs), P, kmax, ptr(R), ptr(t), stream()),
          "kabsch_batched")
    _count(1)
    return R.view(P, 3, 3), t


def get_rigid_transform_3D(reference, target):
    """Same signature as utils/transformation.py:178 ([k,3] arrays -> R [3,3], t [3]); runs K9 on cuda:0."""
    ref = torch.as_tensor(np.asarray(reference, dtype=np.float64)).cuda().unsqueeze(0)
    tgt = torch.as_tensor(np.asarray(target, dtype=np.float64)).cuda().unsqueeze(0)
    if ref.shape != tgt.shape or ref.shape[2] != 3:
        raise Exception(f"matrix A is not 3xN, it is {ref.shape[2]}x{ref.shape[1]}")
    R, t = rigid_transform_3D_batched(ref, tgt)
    return R[0].cpu().numpy(), t[0].cpu().numpy()
