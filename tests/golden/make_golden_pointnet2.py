"""Generate tests/golden/reference_pointnet2.npz by running the reference's OWN model/pointnet2_utils.py and
model/pointnet2.py (pure PyTorch, CPU) on seeded inputs.  python tests/golden/make_golden_pointnet2.py   (needs /root/reference)"""
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def ee_like_cloud(rng, n):
    """n points on the faces of a 0.10 x 0.22 x 0.126 m box under a random rotation, centred (predict_key_points
    centres the crop at the origin, app/inference_engine.py:494-495)."""
    dims = np.array([0.10, 0.22, 0.126])
    face = rng.integers(0, 6, n)
    p = rng.random((n, 3)) * dims
    for f in range(6):
        m = face == f
        p[m, f // 2] = dims[f // 2] * (f % 2)
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    w, x, y, z = q
    R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                  [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                  [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
    p = (p - dims / 2) @ R.T + rng.normal(0, 0.0016, (n, 3))
    return p.astype(np.float32)


def main():
    m = types.ModuleType("ipdb")
    m.set_trace = lambda *a, **k: None
    sys.modules["ipdb"] = m
    sys.path.insert(0, REF)
    from model import pointnet2_utils as RU
    from model.pointnet2 import PointNet2SSG
    rng = np.random.default_rng(13)
    B, N = 3, 2048
    xyz = torch.from_numpy(np.stack([ee_like_cloud(rng, N) for _ in range(B)]))
    out = dict(xyz=xyz.numpy())
    torch.manual_seed(7)
    fps = RU.farthest_point_sample(xyz, 512)
    out["fps_idx"] = fps.numpy().astype(np.int32)
    new_xyz = RU.index_points(xyz, fps)
    for r, ns, tag in ((0.1, 32, "r10"), (0.03, 32, "r03"), (0.015, 16, "r015")):
        out["ball_" + tag] = RU.query_ball_point(r, ns, xyz, new_xyz).numpy().astype(np.int32)
        d = RU.square_distance(new_xyz, xyz)
        out["ball_margin_" + tag] = (d - r * r).abs().min(-1)[0].numpy()  # per query: closest approach to the boundary
    d = RU.square_distance(xyz, new_xyz)
    ds, idx = d.sort(dim=-1)
    ds, idx = ds[:, :, :3], idx[:, :, :3]
    rec = 1.0 / (ds + 1e-8)
    out["nn_idx"] = idx.numpy().astype(np.int32)
    out["nn_w"] = (rec / rec.sum(2, keepdim=True)).numpy()
    out["nn_d"] = ds.numpy()
    # whole network: seeded init, eval mode, seeded FPS starts (4 torch.randint draws inside forward)
    torch.manual_seed(13)
    net = PointNet2SSG(num_classes=6, in_channels=6).eval()
    g = torch.Generator().manual_seed(5)
    feats = torch.rand(2, 3, N, generator=g) - 0.5
    inp = torch.cat((xyz[:2].permute(0, 2, 1), feats), dim=1)  # [2, 6, N]: centred xyz + colours
    torch.manual_seed(99)
    with torch.no_grad():
        logits, l4 = net(inp)
    out.update(net_in=inp.numpy(), net_logits=logits.numpy(), net_l4=l4.numpy())
    np.savez_compressed(os.path.join(HERE, "reference_pointnet2.npz"), **out)
    print({k: v.shape for k, v in out.items()})
    for tag in ("r10", "r03", "r015"):
        print(tag, "queries within 1e-6 of the radius:", int((out["ball_margin_" + tag] < 1e-6).sum()))


if __name__ == "__main__":
    main()
