"""Oracle (TEST INFRASTRUCTURE) for the geometry stages: Kabsch, ICP, clustering, per-point heads,
translation. NumPy / SciPy / scikit-learn, float64, one frame at a time like the reference.

Each function cites the reference code it restates. Where the arithmetic lives in an absent third-party
library (Open3D 0.15.2 registration_icp, scikit-learn 0.24.2 AgglomerativeClustering) the published
algorithm is restated (SURVEY.md §8a rows a16, a21 [3P-memory]) — parity for those rows is UNPINNED.
"""
import numpy as np
import torch
from scipy.sparse import coo_matrix
from scipy.sparse.csgraph import connected_components
from scipy.spatial import cKDTree
from scipy.spatial.transform import Rotation


# ---- utils/transformation.py -------------------------------------------------------------------
def quaternion_rotation_matrix(Q, switch_w=True):
    """utils/transformation.py:16-60."""
    Q = np.asarray(Q)
    if switch_w:
        Q = np.insert(Q[:3], 0, Q[-1])
    q0, q1, q2, q3 = Q[0], Q[1], Q[2], Q[3]
    return np.array([[2 * (q0 * q0 + q1 * q1) - 1, 2 * (q1 * q2 - q0 * q3), 2 * (q1 * q3 + q0 * q2)],
                     [2 * (q1 * q2 + q0 * q3), 2 * (q0 * q0 + q2 * q2) - 1, 2 * (q2 * q3 - q0 * q1)],
                     [2 * (q1 * q3 - q0 * q2), 2 * (q2 * q3 + q0 * q1), 2 * (q0 * q0 + q3 * q3) - 1]])


def transformation_matrix(pose, switch_w=False):
    """utils/transformation.py:63-68."""
    pose = np.asarray(pose)
    T = np.eye(4)
    T[:3, :3] = quaternion_rotation_matrix(pose[3:], switch_w=switch_w)
    T[:3, 3] = pose[:3]
    return T


def q_from_matrix(R):
    """utils/transformation.py:80-84 (SciPy xyzw -> wxyz, sign as SciPy returns it)."""
    q = Rotation.from_matrix(np.array(R, copy=True)).as_quat()
    return np.insert(q[:3], 0, q[-1])


def pose_from_matrix(T):
    """utils/transformation.py:87-93."""
    return np.concatenate((T[:3, 3], q_from_matrix(T[:3, :3])))


def rigid_transform_3D(reference, target):
    """utils/transformation.py:178-222: Kabsch with the reflection fix on the last row of Vt."""
    A, B = np.asarray(reference, np.float64).T, np.asarray(target, np.float64).T
    cA, cB = A.mean(axis=1, keepdims=True), B.mean(axis=1, keepdims=True)
    H = (A - cA) @ (B - cB).T
    U, S, Vt = np.linalg.svd(H)
    R = Vt.T @ U.T
    if np.linalg.det(R) < 0:
        Vt[2, :] *= -1
        R = Vt.T @ U.T
    t = -R @ cA + cB
    return R, t.reshape(-1)


def rotation_angle_deg(Ra, Rb):
    c = (np.trace(Ra.T @ Rb) - 1.0) / 2.0
    return float(np.degrees(np.arccos(np.clip(c, -1.0, 1.0))))


# ---- utils/output.py ---------------------------------------------------------------------------
def segmentation_labels(logits):
    """utils/output.py:67-73: arg-max over classes, lowest index on ties (torch.max semantics)."""
    return torch.as_tensor(logits).float().max(1)[1].numpy()


def largest_cluster(points, dist=0.06):
    """utils/output.py:13-28 restated: single-linkage clusters with distance_threshold=dist are the connected
    components of {d(i,j) < dist}; returns the indices of the largest (ties: component of the lowest index)."""
    pts = np.asarray(points, dtype=np.float64)
    n = len(pts)
    if n == 0:
        return np.zeros(0, np.int64)
    pairs = cKDTree(pts).query_pairs(dist * (1 + 1e-9), output_type="ndarray")
    if len(pairs):
        d = np.linalg.norm(pts[pairs[:, 0]] - pts[pairs[:, 1]], axis=1)
        pairs = pairs[d < dist]
    g = coo_matrix((np.ones(len(pairs)), (pairs[:, 0], pairs[:, 1])), shape=(n, n)) if len(pairs) else coo_matrix((n, n))
    _, lab = connected_components(g, directed=False)
    counts = np.bincount(lab)
    best = counts.max()
    first_of = np.full(len(counts), n)
    np.minimum.at(first_of, lab, np.arange(n))
    cand = np.nonzero(counts == best)[0]
    win = cand[np.argmin(first_of[cand])]
    return np.nonzero(lab == win)[0]


def largest_cluster_sklearn(points, dist=0.06):
    """the reference's own call (utils/output.py:15-28) with `metric=` instead of the removed `affinity=`."""
    from sklearn.cluster import AgglomerativeClustering
    lab = AgglomerativeClustering(distance_threshold=dist, n_clusters=None, metric="euclidean",
                                  linkage="single").fit(np.asarray(points)).labels_
    uniq, counts = np.unique(lab, return_counts=True)
    return np.where(lab == uniq[counts.argmax()])[0]


def key_point_predictions(logits, conf_th=0.999):
    """utils/output.py:81-87."""
    sm = torch.as_tensor(logits).float().softmax(1).max(0)
    classes = np.where(sm[0] > conf_th)[0]
    return sm[1][classes].numpy(), classes, sm[0][classes]


def key_point_best(logits):
    """per class (best probability, lowest row reaching it) — the quantities K8a returns."""
    p = torch.as_tensor(logits).float().softmax(1)
    best = p.max(0)[0]
    idx = torch.stack([torch.nonzero(p[:, c] == best[c])[0, 0] for c in range(p.shape[1])])
    return best.numpy(), idx.numpy()


def pred_center(out, coords, ee_r=0.03, q=None, topk=8):
    """utils/output.py:45-64 (ties in the sort resolved to the lowest index)."""
    v = torch.as_tensor(out)[:, 1].float().numpy()
    order = np.lexsort((np.arange(len(v)), -v))[:topk]
    c = np.asarray(coords, dtype=np.float32)[order].mean(axis=0)
    if q is not None:
        qn = np.asarray(q, dtype=np.float64)
        R = quaternion_rotation_matrix(qn / np.linalg.norm(qn), switch_w=False)
        c = c + R @ np.array([-ee_r, 0, 0])
    return c


def translation_magic(ee_raw_points, q, x_offset=-0.015):
    """app/inference_engine.py:459-489 with magic_enabled: float32 rotation, float64 tail."""
    pts = np.asarray(ee_raw_points, dtype=np.float32)
    R = quaternion_rotation_matrix(np.asarray(q, dtype=np.float32), switch_w=False)
    p = (R.T @ pts.reshape((-1, 3, 1))).reshape((-1, 3))
    off = (p.max(axis=0) + p.min(axis=0)) / 2
    min_z = (p - off).min(axis=0)[2]
    return R @ (np.array([x_offset, 0.0, min_z]) + off)


# ---- utils/icp.py ------------------------------------------------------------------------------
def icp_point_to_point(source, target, init_T, max_corr=0.1, max_iter=30, rel_fitness=1e-6, rel_rmse=1e-6):
    """Open3D 0.15.2 registration_icp(source, target, max_corr, init, PointToPoint) as called at
    utils/icp.py:65-71, restated [3P-memory]: hybrid (radius, 1-NN) KD-tree correspondences, umeyama without
    scaling, stop when |d fitness| and |d rmse| both fall below 1e-6, at most 30 iterations.
    Returns (T 4x4, fitness, inlier_rmse, iterations)."""
    src = np.asarray(source, dtype=np.float64)
    tgt = np.asarray(target, dtype=np.float64)
    tree = cKDTree(tgt)
    T = np.array(init_T, dtype=np.float64)

    def evaluate(p):
        d, j = tree.query(p, k=1)
        ok = d < max_corr
        n = int(ok.sum())
        if n == 0:
            return ok, j, 0.0, 0.0
        return ok, j, n / len(p), float(np.sqrt((d[ok] ** 2).sum() / n))

    p = src @ T[:3, :3].T + T[:3, 3]
    ok, j, fit, rmse = evaluate(p)
    it = 0
    for it in range(1, max_iter + 1):
        if ok.any():
            R, t = rigid_transform_3D(p[ok], tgt[j[ok]])
            U = np.eye(4)
            U[:3, :3], U[:3, 3] = R, t
            T = U @ T
            p = p @ R.T + t          # Open3D transforms the source cloud in place by each update
        ok, j, fit2, rmse2 = evaluate(p)
        done = abs(fit - fit2) < rel_fitness and abs(rmse - rmse2) < rel_rmse
        fit, rmse = fit2, rmse2
        if done:
            break
    return T, fit, rmse, it


# ---- ingest (utils/ros_utils.py:142-167, app/freenect_data_engine.py:81, utils/preprocess.py:20-37, utils/data.py:58-75)
def roi_mask(points, min_x=-500, max_x=500, min_y=-500, max_y=500, min_z=-500, max_z=500, offset=0.0):
    """utils/data.py:58-75: strict box test (and x > -500)."""
    p = np.asarray(points)
    # the reference compares the (float32) array with Python scalars: the comparison runs in the array's dtype
    lo = (np.array([min_x, min_y, min_z], dtype=np.float64) - offset).astype(p.dtype)
    hi = (np.array([max_x, max_y, max_z], dtype=np.float64) + offset).astype(p.dtype)
    return (p[:, 0] > -500) & np.all(p < hi, axis=1) & np.all(p > lo, axis=1)


def ingest_records(rec, roi=None):
    """[n,4] f32 records (x, y, z, PCL-packed rgb) of ONE frame -> (points f32 [n',3], rgb f32 [n',3] in
    [-0.5, 0.5], kept input indices): non-finite points dropped, rgb split, / 255, - 0.5, ROI mask."""
    rec = np.asarray(rec, dtype=np.float32)
    ok = np.isfinite(rec[:, 0]) & np.isfinite(rec[:, 1]) & np.isfinite(rec[:, 2])
    bits = rec[:, 3].copy().view(np.uint32)
    rgb = np.stack(((bits >> 16) & 255, (bits >> 8) & 255, bits & 255), axis=1).astype(np.float64)
    pts = rec[:, :3]
    with np.errstate(invalid="ignore"):
        ok &= roi_mask(pts, *(roi if roi is not None else ()))
    rgbn = (rgb / 255.0 - 0.5).astype(np.float32)
    keep = np.nonzero(ok)[0]
    return pts[keep], rgbn[keep], keep
