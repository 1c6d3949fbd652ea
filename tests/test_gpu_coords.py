"""K1-K3 parity: voxel coordinates, inverse maps, stride maps, kernel maps bit-exact; UNWEIGHTED_AVERAGE voxel
features equal to the oracle's order-independent fixed-point definition and within 1 ulp of the float64 mean."""
import numpy as np
import pytest
import torch

import oracle.MinkowskiEngine as OME
from gpu_util import kinect_like_cloud

pytestmark = pytest.mark.gpu


def _cuda_me():
    import MinkowskiEngine as ME
    return ME


def _fields(points_list, feats_list, scale):
    ME = _cuda_me()
    co = OME.utils.batched_coordinates([torch.from_numpy(p) * scale for p in points_list], dtype=torch.float32)
    fe = torch.from_numpy(np.concatenate(feats_list))
    of = OME.TensorField(features=fe, coordinates=co)
    cf = ME.TensorField(features=fe, coordinates=co, device="cuda")
    return of, cf


@pytest.mark.parametrize("n,scale,batches", [(20000, 100.0, 1), (50000, 200.0, 3), (3000, 20.0, 2)])
def test_voxelize_bit_exact(n, scale, batches):
    pts = [kinect_like_cloud(n, 10 + b) - np.float32(b == 1) * 2 for b in range(batches)]  # batch 1 has negatives
    feats = [np.random.default_rng(b).random((n, 3)).astype(np.float32) - 0.5 for b in range(batches)]
    of, cf = _fields(pts, feats, scale)
    os_, cs = of.sparse(), cf.sparse()
    assert torch.equal(cs.C.cpu(), os_.C)                                   # coordinates, first-occurrence order
    assert torch.equal(cf.inverse_mapping.cpu().long(), of.inverse_mapping)  # inverse map
    # voxel mean: the kernel's order-independent 2^32 fixed-point sum, which the oracle restates (so this equality is
    # agreement with that definition, not with ME's float arithmetic) ...
    assert torch.equal(cs.F.cpu(), os_.F)
    assert torch.equal(cs.slice(cf).F.cpu(), os_.slice(of).F)
    # ... and, independently, within 1 ulp of the float64 mean of the member points rounded to fp32
    inv = of.inverse_mapping.numpy()
    fe = np.concatenate(feats).astype(np.float64)
    V = os_.F.shape[0]
    sums = np.zeros((V, fe.shape[1]))
    np.add.at(sums, inv, fe)
    mean64 = sums / np.bincount(inv, minlength=V)[:, None]
    got = cs.F.cpu().numpy()
    ulp = np.spacing(np.abs(mean64).astype(np.float32)).astype(np.float64)
    assert np.all(np.abs(got.astype(np.float64) - mean64) <= ulp), "voxel mean further than 1 ulp from the fp64 mean"


def test_voxelize_edge_cases():
    ME = _cuda_me()
    # all points in one voxel; duplicates; a single point
    for pts in (np.full((1000, 3), 0.3, np.float32), np.array([[-0.5, 0.2, 0.9]], np.float32),
                np.repeat(np.array([[1.5, 2.5, 3.5], [-1.5, -2.5, -3.5]], np.float32), 7, axis=0)):
        fe = torch.arange(len(pts) * 3, dtype=torch.float32).view(-1, 3) / 100
        co = OME.utils.batched_coordinates([torch.from_numpy(pts)], dtype=torch.float32)
        o = OME.TensorField(features=fe, coordinates=co).sparse()
        c = ME.TensorField(features=fe, coordinates=co, device="cuda").sparse()
        assert torch.equal(c.C.cpu(), o.C) and torch.equal(c.F.cpu(), o.F)
    # exact voxel boundaries and negative zero
    pts = np.array([[0.0, -0.0, 1.0], [-1.0, -1e-7, 0.999999], [2.0, 2.0, 2.0], [1.9999999, 2.0, 2.0]], np.float32)
    co = OME.utils.batched_coordinates([torch.from_numpy(pts)], dtype=torch.float32)
    fe = torch.ones(4, 1)
    o = OME.TensorField(features=fe, coordinates=co).sparse()
    c = ME.TensorField(features=fe, coordinates=co, device="cuda").sparse()
    assert torch.equal(c.C.cpu(), o.C)
    # out-of-range coordinate is an error, not a wrong answer
    far = torch.tensor([[0.0, 2e5, 0.0, 0.0]])
    with pytest.raises(ME.B2MEError):
        ME.TensorField(features=torch.ones(1, 1), coordinates=far, device="cuda").sparse()


def test_sparse_tensor_int_coords_first_wins():
    ME = _cuda_me()
    g = torch.Generator().manual_seed(0)
    q = torch.randint(-6, 7, (5000, 3), generator=g)
    co = OME.utils.batched_coordinates([q])
    fe = torch.rand(5000, 4, generator=g)
    o = OME.SparseTensor(fe, co)
    c = ME.SparseTensor(fe, co, device="cuda")
    assert torch.equal(c.C.cpu(), o.C) and torch.equal(c.F.cpu(), o.F)


def test_k3_block_path_equals_direct_path_and_oracle():
    """K3 through 4 x 4 x 4 blocks (b2me_block_rows + b2me_kernel_map_k3_blocks: the path of the large maps) against
    the direct probe kernel and the oracle, at three tensor strides, with negative coordinates and two frames; the
    K3b by-products (row masks, per-offset counts) give the same sort keys and tile masks as the nbr-table kernels."""
    import MinkowskiEngine as ME
    from MinkowskiEngine._lib import lib, ptr, stream, check
    pts = [kinect_like_cloud(60000, 31), kinect_like_cloud(45000, 32) - np.float32(1.5)]
    feats = [np.zeros((len(p), 1), np.float32) for p in pts]
    maps = {}
    for mode, thr in (("blocks", 1), ("direct", 1 << 40)):
        ME.set_k3_block_min_rows(thr)
        try:
            of, cf = _fields(pts, feats, 150.0)
            os_, cs = of.sparse(), cf.sparse()
            om, cm = os_.coordinate_manager, cs.coordinate_manager
            ok, ck = os_.coordinate_map_key, cs.coordinate_map_key
            for level in range(3):
                nbr_c = cm.kernel_map_k3(ck)
                lv = cm.level(ck)
                maps[(mode, level)] = (nbr_c.cpu(), lv.masks_k3[0].cpu(), lv.masks_k3[1].cpu()[:27])
                if mode == "blocks":
                    assert np.array_equal(nbr_c.cpu().numpy().astype(np.int64), om.kernel_map_k3(ok)), f"level {level}"
                    V = lv.V
                    # row masks / counts are what the nbr table says
                    pres = (nbr_c >= 0)
                    want = (pres.long() << torch.arange(27, device="cuda")).sum(1).int()
                    assert torch.equal(lv.masks_k3[0], want)
                    assert torch.equal(lv.masks_k3[1][:27].long(), pres.sum(0))
                    # keys and tile masks from the row masks == from the nbr table
                    ws = torch.empty((128,), dtype=torch.uint8, device="cuda")
                    k_old = torch.empty((V,), dtype=torch.int32, device="cuda")
                    check(lib.b2me_mask_sort_keys(ptr(nbr_c), V, 27, ptr(k_old), ptr(ws), ws.numel(), stream()))
                    k_new = torch.empty((V,), dtype=torch.int32, device="cuda")
                    check(lib.b2me_mask_sort_keys_rows(ptr(lv.masks_k3[0]), ptr(lv.masks_k3[1]), V, 27, ptr(k_new), stream()))
                    assert torch.equal(k_old, k_new)
                    perm = torch.sort(k_new)[1].int()
                    assert torch.equal(ME.tile_masks(nbr_c, perm, V, 27), ME.tile_masks(nbr_c, perm, V, 27, row_masks=lv.masks_k3[0]))
                    assert torch.equal(ME.tile_masks(nbr_c, None, V, 27), ME.tile_masks(nbr_c, None, V, 27, row_masks=lv.masks_k3[0]))
                ok, _ = om.stride_down(ok)
                ck, _ = cm.stride_down(ck)
        finally:
            ME.set_k3_block_min_rows(65536)
    for level in range(3):
        for a, b in zip(maps[("blocks", level)], maps[("direct", level)]):
            assert torch.equal(a, b), f"block path and direct path differ at level {level}"


def test_sparse_collate_feeds_sparse_tensor():
    """a4 -> a5, the reference's training / test call pattern (data/alivev2.py:386-396 then test.py:52
    `ME.SparseTensor(feats, coordinates=coords, device=...)`): ragged samples with duplicates collated on the host,
    voxelised on the GPU; coordinates and first-wins features bit-exact vs the oracle."""
    ME = _cuda_me()
    g = torch.Generator().manual_seed(4)
    coords = [torch.randint(-9, 10, (n, 3), generator=g).float() + 0.25 for n in (3000, 0, 1700)]
    feats = [torch.rand(len(c), 3, generator=g) for c in coords]
    labels = [torch.randint(0, 3, (len(c),), generator=g) for c in coords]
    cb, fb, lb = ME.utils.sparse_collate(coords, feats, labels, dtype=torch.float32)
    ocb, ofb, olb = OME.utils.sparse_collate(coords, feats, labels, dtype=torch.float32)
    assert torch.equal(cb, ocb) and torch.equal(fb, ofb) and torch.equal(lb, olb)
    o = OME.SparseTensor(ofb, ocb)
    c = ME.SparseTensor(fb, cb, device="cuda")
    assert torch.equal(c.C.cpu(), o.C) and torch.equal(c.F.cpu(), o.F)
    assert set(c.C[:, 0].cpu().tolist()) == {0, 2}        # the empty sample keeps its batch index free


def test_stride_and_kernel_maps_bit_exact():
    ME = _cuda_me()
    pts = [kinect_like_cloud(40000, 21), kinect_like_cloud(30000, 22) - np.float32(1.0)]
    feats = [np.zeros((len(p), 1), np.float32) for p in pts]
    of, cf = _fields(pts, feats, 100.0)
    os_, cs = of.sparse(), cf.sparse()
    om, cm = os_.coordinate_manager, cs.coordinate_manager
    ok, ck = os_.coordinate_map_key, cs.coordinate_map_key
    for level in range(4):
        nbr_o = om.kernel_map_k3(ok)
        nbr_c = cm.kernel_map_k3(ck)
        assert np.array_equal(nbr_c.cpu().numpy().astype(np.int64), nbr_o), f"k3 map, level {level}"
        ok2, rec_o = om.stride_down(ok)
        ck2, rec_c = cm.stride_down(ck)
        assert np.array_equal(cm.coordinates(ck2).cpu().numpy(), om.levels[ok2].coords), f"stride coords {level}"
        assert np.array_equal(rec_c["in2out"].cpu().numpy().astype(np.int64), rec_o["in2out"])
        assert np.array_equal(rec_c["koff"].cpu().numpy().astype(np.int64), rec_o["koff"])
        assert np.array_equal(rec_c["nbr_down"].cpu().numpy().astype(np.int64), rec_o["nbr_down"])
        assert np.array_equal(rec_c["nbr_up"].cpu().numpy().astype(np.int64), rec_o["nbr_up"])
        ok, ck = ok2, ck2
    # kernel-map properties that hold at any size: centre = identity, symmetry nbr[nbr[o,k], 26-k] == o
    nbr = cm.kernel_map_k3(cs.coordinate_map_key).long()
    V = nbr.shape[0]
    assert torch.equal(nbr[:, 13], torch.arange(V, device="cuda"))
    for k in (0, 5, 12, 20):
        rows = torch.nonzero(nbr[:, k] >= 0).flatten()
        assert torch.equal(nbr[nbr[rows, k], 26 - k], rows)


def test_sparse_quantize_matches_oracle():
    ME = _cuda_me()
    rng = np.random.default_rng(3)
    pts = (rng.random((20000, 3)) * 2 - 1).astype(np.float32)
    feats = rng.random((20000, 3)).astype(np.float32)
    labels = rng.integers(0, 3, 20000).astype(np.int32)
    a = OME.utils.sparse_quantize(pts, feats, labels, quantization_size=0.01, ignore_label=-100)
    b = ME.utils.sparse_quantize(pts, feats, labels, quantization_size=0.01, ignore_label=-100)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    ia, va = OME.utils.sparse_quantize(pts, return_maps_only=True, return_inverse=True, quantization_size=0.05)
    ib, vb = ME.utils.sparse_quantize(pts, return_maps_only=True, return_inverse=True, quantization_size=0.05)
    assert np.array_equal(ia, ib) and np.array_equal(va, vb)


def test_full_size_properties():
    """BASELINE config-2 frame size (~300k points, 5 mm voxels): size-independent properties."""
    ME = _cuda_me()
    from b200calib.synthetic import make_frame
    f = make_frame(13)
    pts = torch.from_numpy(f["points"])
    co = OME.utils.batched_coordinates([pts * 200.0], dtype=torch.float32)
    fld = ME.TensorField(features=torch.from_numpy(f["rgb"]) - 0.5, coordinates=co, device="cuda")
    st = fld.sparse()
    C, inv = st.C, fld.inverse_mapping.long()
    assert torch.equal(C[inv][:, 1:], torch.floor(co[:, 1:].cuda()).int())           # every point lands in its voxel
    assert torch.unique(C, dim=0).shape[0] == C.shape[0]                             # rows are unique
    first = torch.full((C.shape[0],), 1 << 40, dtype=torch.long, device="cuda")
    first.scatter_reduce_(0, inv, torch.arange(len(inv), device="cuda"), reduce="amin")
    assert bool((first[1:] > first[:-1]).all())                                       # first-occurrence order
    cnt = torch.bincount(inv, minlength=C.shape[0])
    mean = torch.zeros_like(st.F, dtype=torch.float64).index_add_(0, inv, fld.F.double()) / cnt.unsqueeze(1)
    assert torch.allclose(st.F.double(), mean, atol=1e-6)
    st2 = ME.TensorField(features=fld.F, coordinates=co, device="cuda").sparse()      # idempotent / deterministic
    assert torch.equal(st2.C, C) and torch.equal(st2.F, st.F)


def test_ingest_clouds_matches_oracle_and_reference_roi(golden):
    """fused on-device ingest of PointCloud2-style records against the NumPy restatement of the
    reference's host path, and its ROI test against the reference's own get_roi_mask outputs (golden)."""
    from b200calib.ingest import ingest_clouds, pack_xyzrgb
    from oracle import geometry as G
    rng = np.random.default_rng(9)
    frames, recs = [], []
    for f in range(4):
        n = int(rng.integers(1, 5000)) if f != 2 else 0            # frame 2 is empty
        pts = rng.normal(0, 1.0, (n, 3)).astype(np.float32)
        if n:
            pts[rng.random(n) < 0.05] = np.nan                      # invalid depth pixels
            pts[rng.random(n) < 0.01, 1] = np.inf
        rec = pack_xyzrgb(pts, rng.integers(0, 256, (n, 3)))
        frames.append(rec)
    offs = np.concatenate(([0], np.cumsum([len(r) for r in frames]))).astype(np.int32)
    allrec = torch.from_numpy(np.concatenate(frames)).cuda()
    roi = (-0.8, 0.75, -1.2, 0.9, -0.75, 2.0)
    for r in (None, roi):
        pts, rgb, bidx, noffs, src = ingest_clouds(allrec, offs, roi=r, want_source_index=True)
        for f, rec in enumerate(frames):
            op, oc, keep = G.ingest_records(rec, r)
            a, b = int(noffs[f]), int(noffs[f + 1])
            assert b - a == len(op), (f, b - a, len(op))
            assert np.array_equal(pts[a:b].cpu().numpy(), op)
            assert np.array_equal(rgb[a:b].cpu().numpy(), oc), "colour normalisation must be bit-exact"
            assert np.array_equal(src[a:b].cpu().numpy(), keep + offs[f])
            assert np.all(bidx[a:b].cpu().numpy() == f)
    # the ROI test itself against the reference's own function
    rec = pack_xyzrgb(golden["roi_pts"], np.zeros((len(golden["roi_pts"]), 3)))
    _, _, _, _, src = ingest_clouds(torch.from_numpy(rec).cuda(), [0, len(rec)], roi=tuple(golden["roi_limits"]),
                                    want_source_index=True)
    assert np.array_equal(src.cpu().numpy(), np.nonzero(golden["roi_mask"])[0])
