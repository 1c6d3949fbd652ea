"""BASELINE.json configs[3] and [4] as parity-test cases (not bench lines), plus the kernels added for them:

  configs[3]  end-effector ICP refinement on 1 k frames in ONE batched call: the first frames against the oracle
              (1e-4 m / 0.01 deg), all frames through size-independent properties (recovered pose close to the
              ground truth, fitness high, idempotence: a second ICP from the result does not move it)
  configs[4]  voxel-size sweep 0.02 -> 0.005 m: coordinates bit-exact and logits within tolerance at every scale
  K3b/K4b     tile masks against a torch restatement; a dense blob for the cell-level clustering (K7)
"""
import numpy as np
import pytest
import torch

import oracle.MinkowskiEngine as OME
from oracle import geometry as G
from b200calib.models import make_models, randomize_bn_stats
from b200calib.synthetic import make_frame, ee_surface_cloud
from gpu_util import rel_err

pytestmark = pytest.mark.gpu


def test_icp_1k_frames_config3():
    from b200calib.icp import icp_p2p_batched
    rng = np.random.default_rng(3)
    cad = ee_surface_cloud(2048, 13)
    F = 1000
    tg, T0, Tgt = [], [], []
    for f in range(F):
        q = rng.normal(size=4)
        q /= np.linalg.norm(q)
        R = G.quaternion_rotation_matrix(q, switch_w=False)
        t = np.array([rng.uniform(-0.5, 0.5), rng.uniform(-0.5, 0.5), rng.uniform(0.8, 1.5)])
        n = int(rng.integers(600, 1500))
        sel = rng.choice(len(cad), n, replace=False)
        tg.append((cad[sel].astype(np.float64) @ R.T + t + rng.normal(0, 0.0016, (n, 3))).astype(np.float32))
        Tg = np.eye(4)
        Tg[:3, :3], Tg[:3, 3] = R, t
        Tgt.append(Tg)
        ang = rng.uniform(-0.06, 0.06, 3)  # small-angle perturbation + 1 cm offset
        dR = G.quaternion_rotation_matrix(np.array([1.0, *(ang / 2)]) / np.linalg.norm([1.0, *(ang / 2)]), switch_w=False)
        Ti = Tg.copy()
        Ti[:3, :3] = dR @ R
        Ti[:3, 3] += rng.uniform(-0.01, 0.01, 3)
        T0.append(Ti)
    offs = np.concatenate(([0], np.cumsum([len(x) for x in tg]))).astype(np.int32)
    cad_d = torch.from_numpy(cad).cuda()
    tg_d = torch.from_numpy(np.concatenate(tg)).cuda()
    T, stats = icp_p2p_batched(cad_d, tg_d, offs, torch.from_numpy(np.stack(T0)))
    Th, st = T.cpu().numpy(), stats.cpu().numpy()
    # parity on the first frames
    for i in range(6):
        To, fit, rmse, it = G.icp_point_to_point(cad, tg[i], T0[i])
        assert np.linalg.norm(Th[i][:3, 3] - To[:3, 3]) < 1e-4, i
        assert G.rotation_angle_deg(Th[i][:3, :3], To[:3, :3]) < 0.01, i
        assert int(st[i, 2]) == it
    # properties on all 1000
    Tg = np.stack(Tgt)
    terr = np.linalg.norm(Th[:, :3, 3] - Tg[:, :3, 3], axis=1)
    assert np.all(st[:, 0] > 0.95), "fitness"
    assert np.median(terr) < 1e-3 and np.percentile(terr, 99) < 5e-3, (np.median(terr), terr.max())
    assert np.all(st[:, 2] <= 30) and np.all(st[:, 2] >= 1)
    assert np.all(np.abs(np.linalg.det(Th[:, :3, :3]) - 1) < 1e-9)
    # idempotence: restarting from the converged pose moves it by less than the stopping tolerance allows
    T2, st2 = icp_p2p_batched(cad_d, tg_d, offs, T)
    d = np.linalg.norm(T2.cpu().numpy()[:, :3, 3] - Th[:, :3, 3], axis=1)
    assert np.percentile(d, 99) < 2e-4, d.max()
    assert np.all(st2.cpu().numpy()[:, 1] <= st[:, 1] + 1e-6)  # rmse does not get worse


@pytest.mark.parametrize("scale", [50.0, 100.0, 200.0])
def test_voxel_sweep_config4(scale):
    import MinkowskiEngine as ME
    torch.manual_seed(13)
    f = make_frame(17, width=128, height=96)
    pts, rgb = torch.from_numpy(f["points"]), torch.from_numpy(f["rgb"]) - 0.5
    o = randomize_bn_stats(make_models(OME).RobotNetSegmentation(3, num_classes=3)).eval()
    c = make_models(ME).RobotNetSegmentation(3, num_classes=3)
    c.load_state_dict(o.state_dict())
    c = c.cuda().eval()
    co = OME.utils.batched_coordinates([pts * scale], dtype=torch.float32)
    with torch.no_grad():
        oo = o(OME.TensorField(features=rgb, coordinates=co).sparse())
        for dtype, tol in ((torch.float32, 1e-3), (torch.bfloat16, 2e-2)):
            ME.set_compute_dtype(dtype)
            try:
                cc = c(ME.TensorField(features=rgb, coordinates=co, device="cuda").sparse())
                assert torch.equal(cc.C.cpu(), oo.C), f"voxel coordinates differ at scale {scale}"
                assert rel_err(cc.F.float(), oo.F) < tol, (scale, dtype)
            finally:
                ME.set_compute_dtype(torch.float32)


def test_tile_masks_and_sorted_perm():
    import MinkowskiEngine as ME
    g = torch.Generator().manual_seed(5)
    for V, K in ((1, 27), (255, 8), (257, 27), (5000, 27), (70000, 8)):
        nbr = torch.randint(-3, V, (V, K), generator=g).clamp(min=-1).int()
        nbr[torch.rand(V, K, generator=g) < 0.6] = -1
        nbr_c = nbr.cuda()
        for perm in (None, torch.randperm(V, generator=g).int().cuda(), ME.mask_sorted_perm(nbr_c, V, K)):
            masks = ME.tile_masks(nbr_c, perm, V, K).cpu()
            order = torch.arange(V) if perm is None else perm.long().cpu()
            pres = (nbr[order] >= 0)
            pad = (-V) % 256
            pres = torch.cat((pres, torch.zeros(pad, K, dtype=torch.bool))).view(-1, 256, K).any(1)
            want = (pres.long() << torch.arange(K)).sum(1).int()
            assert torch.equal(masks, want), (V, K)
        perm = ME.mask_sorted_perm(nbr_c, V, K)
        assert torch.equal(torch.sort(perm.long())[0].cpu(), torch.arange(V))
        blk = ME.mask_sorted_perm(nbr_c, V, K, block_rows=1024).long().cpu()
        assert torch.equal(torch.sort(blk)[0], torch.arange(V))
        assert torch.all((blk // 1024)[1:] >= (blk // 1024)[:-1]), "block-prefixed keys keep blocks together"


def test_cluster_dense_blob_and_near_threshold_gaps():
    """K7 on the shapes the cell-level algorithm special-cases: a dense EE-like box (thousands of points per cell
    neighbourhood), two slabs whose gap straddles the 0.06 m threshold, and isolated far points."""
    from b200calib import output as O
    rng = np.random.default_rng(4)
    box = ee_surface_cloud(6000, 1).astype(np.float32) + np.array([0.1, -0.2, 1.2], np.float32)
    segs = [box]
    for gap in (0.0590, 0.0599, 0.0601, 0.0650):
        a = rng.random((1500, 3)).astype(np.float32) * np.array([0.1, 0.1, 0.005], np.float32)
        b = rng.random((900, 3)).astype(np.float32) * np.array([0.1, 0.1, 0.005], np.float32)
        b[:, 2] += 0.005 + gap
        segs.append(np.concatenate((a, b, np.array([[3.0, 3.0, 3.0]], np.float32))))
    offs = np.zeros(len(segs) + 1, np.int32)
    offs[1:] = np.cumsum([len(s) for s in segs])
    mask, sizes = O.largest_cluster_mask(torch.from_numpy(np.concatenate(segs)).cuda(), offs, 0.06)
    mask = mask.cpu().numpy().astype(bool)
    for s, p in enumerate(segs):
        got = np.nonzero(mask[offs[s]:offs[s + 1]])[0]
        want = np.sort(G.largest_cluster(p, 0.06))
        assert np.array_equal(got, want), f"segment {s}: {len(got)} vs {len(want)}"
        assert int(sizes[s]) == len(want)


def test_mask_sort_keys_match_definition():
    """K3b keys (one- and two-level) against the NumPy statement of their definition (oracle/mask_sort.py), bit-exact,
    on random maps and on the k3 map of a synthetic frame; on the frame the two-level order needs fewer (tile, offset)
    passes than the one-level order, which needs far fewer than the natural order."""
    import MinkowskiEngine as ME
    from MinkowskiEngine._lib import lib, ptr, stream, check
    from oracle import mask_sort as MS

    def device_keys(nbr_c, two_level):
        V, K = nbr_c.shape
        keys = torch.empty(V, dtype=torch.int32, device="cuda")
        if two_level:
            ws = torch.empty(lib.b2me_mask_sort_keys2_ws_bytes(V), dtype=torch.uint8, device="cuda")
            check(lib.b2me_mask_sort_keys2(ptr(nbr_c), V, K, ptr(keys), ptr(ws), ws.numel(), stream()))
        else:
            ws = torch.empty(128, dtype=torch.uint8, device="cuda")
            check(lib.b2me_mask_sort_keys(ptr(nbr_c), V, K, ptr(keys), ptr(ws), ws.numel(), stream()))
        return keys.cpu().numpy()

    g = torch.Generator().manual_seed(9)
    cases = []
    for V, K, p in ((1, 27, 0.5), (300, 27, 0.7), (20000, 27, 0.72), (5000, 8, 0.6)):
        nbr = torch.randint(0, V, (V, K), generator=g).int()
        # offsets with very different frequencies, some never / always present
        drop = torch.rand(V, K, generator=g) < (p * torch.linspace(0.2, 1.3, K)).clamp(max=1.0)
        nbr[drop] = -1
        nbr[:, K // 2] = torch.arange(V, dtype=torch.int32)
        cases.append(nbr)
    f = make_frame(311, width=320, height=240)
    fld = ME.TensorField(features=torch.from_numpy(f["rgb"]).cuda(),
                         coordinates=ME.utils.batched_coordinates([torch.from_numpy(f["points"]) * 200.0],
                                                                  dtype=torch.float32), device="cuda")
    sp = fld.sparse()
    frame_nbr = sp.coordinate_manager.kernel_map_k3(sp.coordinate_map_key)[:sp.C.shape[0]].cpu()
    cases.append(frame_nbr)
    for nbr in cases:
        n = nbr.numpy()
        assert np.array_equal(device_keys(nbr.cuda(), False), MS.keys_one_level(n)), nbr.shape
        assert np.array_equal(device_keys(nbr.cuda(), True), MS.keys_two_level(n)), nbr.shape
    # Morton tie-break keys (the default order of the k3 maps) on the frame's map, stride 1 and a fake stride 2
    V = sp.C.shape[0]
    for ts in (1, 2):
        k64 = torch.empty(V, dtype=torch.int64, device="cuda")
        ws = torch.empty(128, dtype=torch.uint8, device="cuda")
        check(lib.b2me_mask_sort_keys_morton(ptr(frame_nbr.cuda()), ptr(sp.C), V, 27, ts, ptr(k64), ptr(ws), ws.numel(),
                                             stream()))
        assert np.array_equal(k64.cpu().numpy(), MS.keys_morton(frame_nbr.numpy(), sp.C.cpu().numpy(), ts)), ts
    perm = ME.mask_sorted_perm(frame_nbr.cuda(), V, 27, coords=sp.C, ts=1)   # one-level unless the switch is on
    ME.set_mask_sort_morton(True)
    try:
        perm_m = ME.mask_sorted_perm(frame_nbr.cuda(), V, 27, coords=sp.C, ts=1)
    finally:
        ME.set_mask_sort_morton(False)
    for pm in (perm, perm_m):
        assert torch.equal(torch.sort(pm.long())[0].cpu(), torch.arange(V))
    n = frame_nbr.numpy()
    p_nat = MS.passes(n, np.arange(len(n)))
    p_one = MS.passes(n, np.argsort(MS.keys_one_level(n), kind="stable"))
    p_two = MS.passes(n, np.argsort(MS.keys_two_level(n), kind="stable"))
    assert p_two < p_one < 0.6 * p_nat, (p_nat, p_one, p_two)
    # the tie-break only reorders rows inside a mask group: same passes as the one-level order, up to group boundaries
    p_mor = MS.passes(n, np.argsort(MS.keys_morton(n, sp.C.cpu().numpy(), 1), kind="stable"))
    assert abs(p_mor - p_one) <= 0.05 * p_one, (p_one, p_mor)
