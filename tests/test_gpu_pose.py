"""K9 / K10 parity: batched Kabsch against the reference's own get_rigid_transform_3D (golden vectors) and
batched ICP against the Open3D-loop restatement, tolerance 1e-4 m and 0.01 degrees (north star)."""
import numpy as np
import pytest
import torch

from oracle import geometry as G

pytestmark = pytest.mark.gpu


def test_kabsch_golden(golden):
    from b200calib.transformation import rigid_transform_3D_batched, get_rigid_transform_3D
    ref = torch.from_numpy(golden["kabsch_ref"]).cuda()
    tgt = torch.from_numpy(golden["kabsch_tgt"]).cuda()
    n = torch.from_numpy(golden["kabsch_n"]).cuda()
    R, t = rigid_transform_3D_batched(ref, tgt, n)
    R, t = R.cpu().numpy(), t.cpu().numpy()
    for i in range(len(R)):
        assert np.allclose(R[i], golden["kabsch_R"][i], atol=1e-8), i
        assert np.allclose(t[i], golden["kabsch_t"][i], atol=1e-8), i
        assert abs(np.linalg.det(R[i]) - 1) < 1e-10
    k = int(golden["kabsch_n"][3])
    R1, t1 = get_rigid_transform_3D(golden["kabsch_ref"][3][:k], golden["kabsch_tgt"][3][:k])
    assert np.allclose(R1, golden["kabsch_R"][3], atol=1e-8) and np.allclose(t1, golden["kabsch_t"][3], atol=1e-8)


def _icp_case(cad, rng, n_tgt):
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    R = G.quaternion_rotation_matrix(q, switch_w=False)
    t = np.array([rng.uniform(-0.5, 0.5), rng.uniform(-0.5, 0.5), rng.uniform(0.8, 1.5)])
    vis = cad[cad[:, 0] > 0.005] if (cad[:, 0] > 0.005).sum() > n_tgt // 2 else cad
    sel = rng.choice(len(vis), n_tgt, replace=len(vis) < n_tgt)
    tgt = (vis[sel] @ R.T + t + rng.normal(0, 0.0016, (n_tgt, 3))).astype(np.float32)
    pose = np.concatenate((t, q)) + rng.uniform(-0.03, 0.03, 7)      # playground/play_ee_icp.py:113 style jitter
    pose[3:] /= np.linalg.norm(pose[3:])
    return tgt, G.transformation_matrix(pose)


def test_icp_matches_oracle(cad_points):
    from b200calib.icp import icp_p2p_batched
    rng = np.random.default_rng(13)
    cases = [_icp_case(cad_points, rng, n) for n in (2048, 4096, 700, 8192, 3000, 2500)]
    tg = np.concatenate([c[0] for c in cases])
    offs = np.concatenate(([0], np.cumsum([len(c[0]) for c in cases]))).astype(np.int32)
    T0 = torch.from_numpy(np.stack([c[1] for c in cases]))
    T, stats = icp_p2p_batched(torch.from_numpy(cad_points).cuda(), torch.from_numpy(tg).cuda(), offs, T0)
    T, stats = T.cpu().numpy(), stats.cpu().numpy()
    for i, (tgt, init) in enumerate(cases):
        To, fit, rmse, it = G.icp_point_to_point(cad_points, tgt, init)
        assert np.linalg.norm(T[i][:3, 3] - To[:3, 3]) < 1e-4, (i, T[i][:3, 3], To[:3, 3])
        assert G.rotation_angle_deg(T[i][:3, :3], To[:3, :3]) < 0.01, i
        assert abs(stats[i, 0] - fit) < 1e-3 and abs(stats[i, 1] - rmse) < 1e-5
        assert int(stats[i, 2]) == it


def test_icp_edge_cases(cad_points):
    from b200calib.icp import icp_p2p_batched, get_point2point_matcher
    cad = torch.from_numpy(cad_points).cuda()
    far = torch.full((100, 3), 50.0).cuda()                         # no correspondence within 0.1 m
    T0 = torch.eye(4, dtype=torch.float64).unsqueeze(0)
    T, stats = icp_p2p_batched(cad, far, [0, 100], T0)
    assert torch.allclose(T.cpu()[0], T0[0]) and float(stats[0, 0]) == 0.0 and float(stats[0, 3]) == 0.0
    T, stats = icp_p2p_batched(cad, torch.zeros((0, 3)).cuda(), [0, 0], T0)   # empty target
    assert torch.allclose(T.cpu()[0], T0[0])
    same, _ = icp_p2p_batched(cad, cad.clone(), [0, len(cad_points)], T0)     # identical clouds: identity
    assert torch.allclose(same.cpu()[0], T0[0], atol=1e-9)


def test_icp_cluster_sizes_agree(cad_points):
    """k_icp_persistent with 1, 2, 4 or 8 CTAs per frame (thread-block cluster sharing one frame): same iteration
    count, poses equal to fp64 summation-order noise (the per-CTA partial sums are added in rank order)."""
    from b200calib.icp import icp_p2p_batched
    rng = np.random.default_rng(3)
    cad = cad_points.astype(np.float32)
    tgs, offs, T0 = [], [0], []
    for f in range(5):
        ang = rng.uniform(-0.5, 0.5)
        T = np.eye(4)
        T[:3, :3] = np.array([[np.cos(ang), -np.sin(ang), 0], [np.sin(ang), np.cos(ang), 0], [0, 0, 1]])
        T[:3, 3] = rng.uniform(-0.3, 0.3, 3) + [0, 0, 1.2]
        sel = cad[rng.permutation(len(cad))[:1500 + 500 * f]].astype(np.float64)
        tgs.append((sel @ T[:3, :3].T + T[:3, 3] + rng.normal(0, 0.001, sel.shape)).astype(np.float32))
        offs.append(offs[-1] + len(sel))
        init = T.copy()
        init[:3, 3] += rng.uniform(-0.02, 0.02, 3)
        T0.append(init)
    cad_d, tg_d = torch.from_numpy(cad).cuda(), torch.from_numpy(np.concatenate(tgs)).cuda()
    T0 = torch.from_numpy(np.stack(T0))
    ref_T, ref_st = icp_p2p_batched(cad_d, tg_d, offs, T0, cluster_size=1)
    for cs in (2, 4, 8, 0):
        T, st = icp_p2p_batched(cad_d, tg_d, offs, T0, cluster_size=cs)
        assert torch.equal(st[:, 2], ref_st[:, 2]), f"iteration counts differ at cluster size {cs}"
        assert float((T - ref_T).abs().max()) < 1e-9
        assert float((st[:, :2] - ref_st[:, :2]).abs().max()) < 1e-9
