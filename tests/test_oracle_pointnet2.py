"""oracle/pointnet2.py against outputs of the reference's own model/pointnet2_utils.py (golden vectors)."""
import os

import numpy as np
import torch

from conftest import GOLDEN
from oracle import pointnet2 as OP


def _g():
    return np.load(os.path.join(GOLDEN, "reference_pointnet2.npz"))


def test_fps_ball_three_nn_vs_reference_golden():
    g = _g()
    xyz = torch.from_numpy(g["xyz"])
    fps = OP.fps(xyz, 512, torch.from_numpy(g["fps_idx"][:, 0].astype(np.int64)))
    assert np.array_equal(fps.numpy(), g["fps_idx"])
    new_xyz = torch.gather(xyz, 1, fps[:, :, None].expand(-1, -1, 3))
    for r, ns, tag in ((0.1, 32, "r10"), (0.03, 32, "r03"), (0.015, 16, "r015")):
        assert np.array_equal(OP.ball_query(r, ns, xyz, new_xyz).numpy(), g["ball_" + tag]), tag
    idx, w, d = OP.three_nn(xyz, new_xyz)
    assert np.array_equal(idx.numpy(), g["nn_idx"]) and np.allclose(w.numpy(), g["nn_w"], atol=1e-6)
