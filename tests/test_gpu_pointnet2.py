"""PointNet++ primitives on the GPU against outputs of the reference's own
model/pointnet2_utils.py / model/pointnet2.py (tests/golden/reference_pointnet2.npz) and against the oracle on
random clouds. Index work is bit-exact; a ball-query membership may only differ when the point sits on the radius
to within fp32 rounding of the expanded distance formula."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import pointnet2 as OP

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def g():
    return np.load(os.path.join(GOLDEN, "reference_pointnet2.npz"))


def _ball_check(got, want, xyz, new_xyz, r):
    bad = np.nonzero((got != want).any(-1))
    if len(bad[0]) == 0:
        return 0
    d = OP.pairwise_sqdist(torch.from_numpy(new_xyz), torch.from_numpy(xyz)).numpy()
    for b, s in zip(*bad):
        assert np.abs(d[b, s] - r * r).min() < 2e-8, (b, s, "ball query differs away from the radius")
    return len(bad[0])


def test_fps_ball_three_nn_vs_reference_golden(g):
    from b200calib import pointnet2_utils as U
    xyz = torch.from_numpy(g["xyz"]).cuda()
    fps = U.farthest_point_sample(xyz, 512, start=torch.from_numpy(g["fps_idx"][:, 0].astype(np.int64)))
    assert np.array_equal(fps.cpu().numpy(), g["fps_idx"]), "FPS indices differ from the reference"
    new_xyz = U.index_points(xyz, fps)
    for r, ns, tag in ((0.1, 32, "r10"), (0.03, 32, "r03"), (0.015, 16, "r015")):
        got = U.query_ball_point(r, ns, xyz, new_xyz).cpu().numpy()
        nbad = _ball_check(got, g["ball_" + tag], g["xyz"], new_xyz.cpu().numpy(), r)
        assert nbad <= 2, (tag, nbad)
    idx, w = U.three_nn(xyz, new_xyz)
    idx, w = idx.cpu().numpy(), w.cpu().numpy()
    same = (idx == g["nn_idx"]).all(-1)
    # a different neighbour order is only possible between (near-)equal distances
    d = g["nn_d"]
    for b, n in zip(*np.nonzero(~same)):
        gaps = np.abs(np.diff(d[b, n]))
        assert gaps.min() < 1e-7 or set(idx[b, n]) == set(g["nn_idx"][b, n])
    assert same.mean() > 0.999
    assert np.allclose(w[same], g["nn_w"][same], rtol=1e-4, atol=1e-6)


def test_primitives_vs_oracle_random_and_edge_cases():
    from b200calib import pointnet2_utils as U
    gen = torch.Generator().manual_seed(3)
    for B, N, S, r, ns in ((1, 33, 8, 0.3, 4), (4, 1000, 128, 0.2, 32), (2, 8192, 1024, 0.05, 16), (3, 64, 64, 5.0, 64)):
        xyz = torch.rand(B, N, 3, generator=gen) - 0.5
        start = torch.randint(0, N, (B,), generator=gen)
        want = OP.fps(xyz, S, start)
        got = U.farthest_point_sample(xyz.cuda(), S, start=start).cpu()
        assert torch.equal(got, want), (B, N, S)
        new_xyz = torch.gather(xyz, 1, want[:, :, None].expand(-1, -1, 3))
        gb = U.query_ball_point(r, ns, xyz.cuda(), new_xyz.cuda()).cpu().numpy()
        assert _ball_check(gb, OP.ball_query(r, ns, xyz, new_xyz).numpy(), xyz.numpy(), new_xyz.numpy(), r) <= 2
        gi, gw = U.three_nn(xyz.cuda(), new_xyz.cuda())
        oi, ow, od = OP.three_nn(xyz, new_xyz)
        same = (gi.cpu() == oi).all(-1)
        assert float(same.float().mean()) > 0.995
        assert torch.allclose(gw.cpu()[same], ow[same], rtol=1e-4, atol=1e-6)
    # duplicate points: FPS ties go to the lowest index, like torch.max
    xyz = torch.zeros(1, 40, 3)
    xyz[0, 20:] = 1.0
    got = U.farthest_point_sample(xyz.cuda(), 4, start=torch.tensor([0])).cpu()
    assert torch.equal(got, OP.fps(xyz, 4, torch.tensor([0])))


def test_pointnet2_ssg_matches_reference_model_output(g):
    """the mirror of model/pointnet2.py on the GPU primitives reproduces the logits of the reference's own network
    (same seeded initialisation, same seeded FPS starts); fp32 library GEMMs with TF32 switched off."""
    from b200calib.pointnet2 import PointNet2SSG
    tf32 = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        torch.manual_seed(13)
        net = PointNet2SSG(num_classes=6, in_channels=6).eval().cuda()
        inp = torch.from_numpy(g["net_in"]).cuda()
        torch.manual_seed(99)
        with torch.no_grad():
            logits, l4 = net(inp)
        a, b = logits.cpu().double(), torch.from_numpy(g["net_logits"]).double()
        err = float((a - b).norm() / b.norm())
        print("PointNet2SSG logits rel err vs the reference model:", err)
        assert err < 1e-4
        assert torch.allclose(l4.cpu(), torch.from_numpy(g["net_l4"]), atol=1e-4)
        # key-point head on top (utils/output.py:81-87) picks the same points
        from b200calib import output as O
        from oracle import geometry as G
        idx, cls, pr = O.get_key_point_predictions(logits[0], 0.0)
        oi, oc, opr = G.key_point_predictions(g["net_logits"][0], 0.0)
        assert np.array_equal(cls, oc) and np.allclose(pr.numpy(), np.asarray(opr), atol=1e-5)
        ref_p = torch.from_numpy(g["net_logits"][0]).softmax(1).numpy()
        for c, i, j in zip(cls, idx, oi):   # a random-init net is nearly flat: another point may win by < 1e-5
            assert i == j or abs(ref_p[i, c] - ref_p[j, c]) < 1e-5
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32


def test_pipeline_pointnet2_key_point_branch():
    """the reference's default key-point branch (PointNet2SSG on a uniform sample of every EE crop,
    app/inference_engine.py:511-537) inside the batched pipeline: sampled rows stay inside their crop, the reported
    probability is the soft-max value of the reported row, crops below num_dense_points get no key points, and the
    Kabsch / ICP stages consume the result."""
    import MinkowskiEngine as ME
    from b200calib.models import make_models, randomize_bn_stats
    from b200calib.pipeline import BatchedInferenceEngine, PipelineConfig
    from b200calib.pointnet2 import PointNet2SSG
    from b200calib.synthetic import make_frame, ee_surface_cloud
    torch.manual_seed(3)
    M = make_models(ME)
    seg = randomize_bn_stats(M.RobotNetSegmentation(3, num_classes=3, variant="MinkUNet14A")).cuda().eval()
    rot = randomize_bn_stats(M.RobotNetEncode(3, 7, variant="MinkUNet14A")).cuda().eval()
    kp = PointNet2SSG(num_classes=6, in_channels=6).cuda().eval()
    cfg = PipelineConfig(seg_scale=50.0, rot_scale=100.0, ee_point_counts_threshold=64, num_dense_points=512,
                         kp_conf_threshold=0.0)
    eng = BatchedInferenceEngine(seg, rot, kp, cad_points=torch.from_numpy(ee_surface_cloud(1024, 1)).cuda(), config=cfg)
    frames = [make_frame(50 + i, width=320, height=240) for i in range(3)]
    frames.append(make_frame(60, width=96, height=72))           # EE crop smaller than num_dense_points
    res = eng.predict_batch([(f["points"], f["rgb"]) for f in frames], gt_labels=[f["labels"] for f in frames],
                            kp_conf_threshold=0.0)
    big, sample = eng.last_kp_sample
    assert sample.shape[1] == 512 and len(big) >= 3
    n_kp = 0
    for i, (f, r) in enumerate(zip(frames, res)):
        n_ee = int((f["labels"] == 2).sum())
        if n_ee >= 512:
            assert r.key_points_pose is not None and r.ee_pose is not None, i
            assert abs(np.linalg.norm(r.key_points_pose[3:]) - 1) < 1e-6
            n_kp += 1
        elif r.ee_pose is not None:
            assert r.key_points_pose is None, "a crop below num_dense_points must not produce key points"
    assert n_kp >= 3
