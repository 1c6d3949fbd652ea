"""helpers shared by the -m gpu parity tests (CUDA path vs oracle on the same seeded inputs)."""
import numpy as np
import torch


def rel_err(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def kinect_like_cloud(n, seed, extent=(3.0, 2.0), depth=(1.0, 3.0)):
    """n points on a few noisy surfaces in metres (scan-line-ish order), float32."""
    rng = np.random.default_rng(seed)
    u = np.sort(rng.random(n)) * extent[0] - extent[0] / 2
    v = rng.random(n) * extent[1] - extent[1] / 2
    z = depth[0] + (depth[1] - depth[0]) * (0.5 + 0.3 * np.sin(2 * u) * np.cos(3 * v)) + rng.normal(0, 0.002, n)
    blob = rng.random(n) < 0.2
    z[blob] = 1.2 + 0.1 * rng.random(blob.sum())
    return np.stack((u, v, z), 1).astype(np.float32)


def copy_state(src_model, dst_model):
    dst_model.load_state_dict(src_model.state_dict())
    return dst_model
