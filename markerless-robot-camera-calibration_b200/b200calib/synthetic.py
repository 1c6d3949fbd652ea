"""Synthetic Kinect-1-shaped frames: the reference's dataset and checkpoints are unpublished,
so every measurement uses seeded synthetic clouds of the same shape.

A 640x480 pin-hole depth render (Kinect-1 intrinsics of utils/aruco.py:16-22) of a room (back wall + floor +
side wall), a robot "arm" (a chain of 5 cm-radius sphere-swept links) and an end-effector box of
0.10 x 0.22 x 0.126 m (utils/data.py:79-86) at a random pose 1-1.5 m from the camera, with 1.6 mm depth noise
(utils/augmentation.py:49) and 2-5 % of the pixels dropped as invalid. Frames follow the pickle schema of
README.md:52-63: points [N,3] f32, rgb [N,3] f32 in [0,1], labels [N] (0 bg, 1 arm, 2 ee), pose x,y,z,qx,qy,qz,qw.
"""
import numpy as np

FX, FY, CX, CY = 520.34, 513.83, 323.06, 263.50
W, H = 640, 480
EE_DIM = np.array([0.10, 0.22, 0.126])

# app/inference_engine.py:128-137 (EE frame key points used by predict_pose_from_kp)
REFERENCE_KEY_POINTS = np.array([
    [0.01982731, 0.08085986, 0.00321919],
    [0.02171595, -0.08986182, 0.00388430],
    [0.01288678, 0.09103118, 0.06127814],
    [0.02079032, -0.09790908, 0.05609143],
    [-0.00185802, 0.04654205, 0.11564558],
    [0.00241113, -0.04262756, 0.11564558],
])


def _random_rotation(rng):
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    w, x, y, z = q
    R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                  [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                  [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
    return R, q  # q = w,x,y,z


def make_frame(seed=13, width=W, height=H, wide=False):
    """-> dict(points, rgb, labels, pose (x,y,z,qx,qy,qz,qw), ee_pose_wxyz (x,y,z,qw,qx,qy,qz))."""
    rng = np.random.default_rng(seed)
    sx, sy = width / W, height / H
    u, v = np.meshgrid(np.arange(width, dtype=np.float64), np.arange(height, dtype=np.float64))
    d = np.stack(((u - CX * sx) / (FX * sx), (v - CY * sy) / (FY * sy), np.ones_like(u)), axis=-1).reshape(-1, 3)
    n = len(d)
    depth = np.full(n, np.inf)
    label = np.zeros(n, dtype=np.int64)

    def hit(t, lab):
        nonlocal depth, label
        better = (t > 0.3) & (t < depth)
        depth = np.where(better, t, depth)
        label = np.where(better, lab, label)

    # room: back wall z = zw, floor y = yf (camera y points down), side wall x = xw
    zw = rng.uniform(2.4, 3.0) * (1.5 if wide else 1.0)
    yf = rng.uniform(0.7, 1.0)
    xw = rng.uniform(1.2, 1.6)
    hit(zw / d[:, 2], 0)
    with np.errstate(divide="ignore", invalid="ignore"):
        hit(np.where(d[:, 1] > 1e-6, yf / d[:, 1], np.inf), 0)
        hit(np.where(d[:, 0] > 1e-6, xw / d[:, 0], np.inf), 0)

    # end-effector pose and arm joints (camera frame)
    ee_t = np.array([rng.uniform(-0.3, 0.3), rng.uniform(-0.25, 0.15), rng.uniform(1.0, 1.5)])
    R, q_wxyz = _random_rotation(rng)
    base = np.array([rng.uniform(-0.5, 0.5), yf, rng.uniform(1.6, 2.2)])
    wrist = ee_t - R[:, 2] * 0.02
    mid1 = base + np.array([0.0, -0.45, 0.0]) + rng.normal(0, 0.05, 3)
    mid2 = (mid1 + wrist) / 2 + rng.normal(0, 0.08, 3)
    joints = [base, mid1, mid2, wrist]

    # arm: spheres swept along the links (5 cm radius)
    centers = []
    for a, b in zip(joints[:-1], joints[1:]):
        m = max(2, int(np.linalg.norm(b - a) / 0.02))
        centers.append(a + (b - a) * np.linspace(0, 1, m)[:, None])
    centers = np.concatenate(centers)
    r = 0.05
    dd = (d * d).sum(1)
    for c in centers:
        b = d @ c
        disc = b * b - dd * (c @ c - r * r)
        t = np.where(disc > 0, (b - np.sqrt(np.maximum(disc, 0))) / dd, np.inf)
        hit(t, 1)

    # end-effector: oriented box, slab test in the EE frame (box spans [0,dx] x [-dy/2,dy/2] x [0,dz])
    lo = np.array([0.0, -EE_DIM[1] / 2, 0.0])
    hi = np.array([EE_DIM[0], EE_DIM[1] / 2, EE_DIM[2]])
    o_l = R.T @ (-ee_t)
    d_l = d @ R
    with np.errstate(divide="ignore", invalid="ignore"):
        t1 = (lo - o_l) / d_l
        t2 = (hi - o_l) / d_l
    tmin = np.nanmax(np.minimum(t1, t2), axis=1)
    tmax = np.nanmin(np.maximum(t1, t2), axis=1)
    hit(np.where((tmax >= tmin) & (tmax > 0), tmin, np.inf), 2)

    depth = depth * (1.0 + 0.0)  # along-ray parameter with d_z = 1 -> z = depth
    depth = depth + rng.normal(0, 0.0016, n)
    keep = np.isfinite(depth) & (rng.random(n) > rng.uniform(0.02, 0.05))
    pts = (d * depth[:, None])[keep].astype(np.float32)
    labels = label[keep]
    rgb = rng.random((len(pts), 3)).astype(np.float32)
    pose_xyzw = np.concatenate((ee_t, q_wxyz[1:], q_wxyz[:1]))
    return dict(points=pts, rgb=rgb, labels=labels, instance_labels=labels.copy(), pose=pose_xyzw.astype(np.float32),
                ee_pose_wxyz=np.concatenate((ee_t, q_wxyz)).astype(np.float64),
                joint_angles=np.zeros(9, dtype=np.float32))


def make_frames(n, seed=13, **kw):
    return [make_frame(seed * 1000 + i, **kw) for i in range(n)]


def ee_surface_cloud(n=4096, seed=13):
    """seeded points on the surface of the EE box (EE frame) — a stand-in CAD cloud for ICP runs on machines
    where app/hand_files/ is not present."""
    rng = np.random.default_rng(seed)
    dx, dy, dz = EE_DIM
    areas = np.array([dy * dz, dy * dz, dx * dz, dx * dz, dx * dy, dx * dy])
    face = rng.choice(6, size=n, p=areas / areas.sum())
    a, b = rng.random(n), rng.random(n)
    p = np.zeros((n, 3))
    for f in range(6):
        m = face == f
        if f < 2:
            p[m] = np.stack((np.full(m.sum(), dx * (f % 2)), (a[m] - 0.5) * dy, b[m] * dz), 1)
        elif f < 4:
            p[m] = np.stack((a[m] * dx, np.full(m.sum(), (f % 2 - 0.5) * dy), b[m] * dz), 1)
        else:
            p[m] = np.stack((a[m] * dx, (b[m] - 0.5) * dy, np.full(m.sum(), dz * (f % 2))), 1)
    return p.astype(np.float32)
