"""K5-K8 parity: pooling, key-point / vote reductions, translation, largest EE cluster vs the oracle."""
import numpy as np
import pytest
import torch

import oracle.MinkowskiEngine as OME
from oracle import geometry as G

pytestmark = pytest.mark.gpu


def test_largest_cluster_matches_oracle_and_sklearn():
    from b200calib import output as O
    rng = np.random.default_rng(0)
    segs = []
    for s in range(5):
        blobs = [rng.normal(rng.uniform(-0.5, 0.5, 3), 0.03, (int(rng.integers(50, 900)), 3)) for _ in range(4)]
        segs.append(np.concatenate(blobs).astype(np.float32))
    segs.append(rng.random((3, 3)).astype(np.float32) * 5)      # tiny segment, all isolated
    segs.append(np.zeros((0, 3), np.float32))                   # empty segment
    offs = np.zeros(len(segs) + 1, np.int32)
    offs[1:] = np.cumsum([len(s) for s in segs])
    allp = torch.from_numpy(np.concatenate(segs)).cuda()
    mask, sizes = O.largest_cluster_mask(allp, offs, 0.06)
    mask = mask.cpu().numpy().astype(bool)
    for s, p in enumerate(segs):
        got = np.nonzero(mask[offs[s]:offs[s + 1]])[0]
        want = G.largest_cluster(p, 0.06)
        assert np.array_equal(got, np.sort(want)), f"segment {s}"
        assert int(sizes[s]) == len(want)
        if 10 < len(p) < 3000:
            assert np.array_equal(got, np.sort(G.largest_cluster_sklearn(p, 0.06)))
    idx = O.ClusterUtil().get_largest_cluster(segs[0])
    assert np.array_equal(idx, np.sort(G.largest_cluster(segs[0])))


def test_keypoint_vote_translation():
    from b200calib import output as O
    g = torch.Generator().manual_seed(1)
    sizes = [700, 2048, 1, 0, 333]
    offs = np.concatenate(([0], np.cumsum(sizes))).astype(np.int32)
    n = int(offs[-1])
    logits = torch.randn(n, 6, generator=g) * 3
    pts = torch.rand(n, 3, generator=g)
    bp, bi = O.key_point_predictions_batched(logits.cuda(), offs)
    vc = O.vote_centers_batched(logits.cuda(), pts.cuda(), offs, col=1, topk=8)
    quat = torch.nn.functional.normalize(torch.randn(len(sizes), 4, generator=g), dim=1)
    tr = O.translation_magic_batched(pts.cuda(), offs, quat.cuda())
    for s in range(len(sizes)):
        a, b = offs[s], offs[s + 1]
        if b == a:
            continue
        p, i = G.key_point_best(logits[a:b])
        assert np.allclose(bp[s].cpu().numpy(), p, atol=2e-6)
        same = (bi[s].cpu().numpy() - a) == i
        # a different row may only win when its probability is within rounding of the best
        probs = torch.softmax(logits[a:b], 1).numpy()
        for c in np.nonzero(~same)[0]:
            assert abs(probs[bi[s, c].item() - a, c] - p[c]) < 2e-6
        want = G.pred_center(logits[a:b], pts[a:b].numpy())
        assert np.allclose(vc[s].cpu().numpy(), want, atol=1e-6)
        wt = G.translation_magic(pts[a:b].numpy(), quat[s].numpy())
        assert np.allclose(tr[s].cpu().numpy(), wt, atol=1e-5)     # tolerance 1e-4 m (north star); fp32 inside
    idx, classes, probs = O.get_key_point_predictions(logits[:700], conf_th=0.5)
    oi, oc, op = G.key_point_predictions(logits[:700], conf_th=0.5)
    assert np.array_equal(classes, oc) and np.array_equal(idx, oi)


def test_global_pool_and_slice():
    import MinkowskiEngine as ME
    g = torch.Generator().manual_seed(2)
    pts = [torch.rand(n, 3, generator=g) * 30 for n in (4000, 10, 2500)]
    fe = torch.randn(6510, 16, generator=g)
    co = OME.utils.batched_coordinates(pts, dtype=torch.float32)
    o = OME.TensorField(features=fe, coordinates=co).sparse()
    c = ME.TensorField(features=fe, coordinates=co, device="cuda").sparse()
    for mod in ("MinkowskiGlobalAvgPooling", "MinkowskiGlobalMaxPooling"):
        po = getattr(OME, mod)()(o).F
        pc = getattr(ME, mod)()(c).F.cpu()
        assert po.shape == pc.shape == (3, 16)
        assert torch.allclose(pc, po, atol=1e-5)
    bn = ME.MinkowskiBatchNorm(16).cuda().eval()
    obn = OME.MinkowskiBatchNorm(16).eval()
    with torch.no_grad():
        bn.bn.running_mean.normal_(); bn.bn.running_var.uniform_(0.5, 2); bn.bn.weight.normal_(); bn.bn.bias.normal_()
    obn.load_state_dict({k: v.cpu() for k, v in bn.state_dict().items()})
    relu_c, relu_o = ME.MinkowskiLeakyReLU(), OME.MinkowskiLeakyReLU()
    with torch.no_grad():
        yc = relu_c(bn(c)) + c          # stand-alone affine+act kernel, then an un-fused add
        yo = relu_o(obn(o)) + o
    assert torch.allclose(yc.F.cpu(), yo.F, atol=1e-5)
    assert torch.allclose(ME.cat(yc, c).F.cpu(), OME.cat(yo, o).F, atol=1e-5)


def test_heads_vs_reference_golden(golden):
    """K5 / K8 against outputs of the reference's OWN utils/output.py functions (golden vectors)."""
    from b200calib import output as O
    lg = torch.from_numpy(golden["kp_logits"])
    for th, tag in ((0.75, "75"), (0.999, "999")):
        idx, cls, pr = O.get_key_point_predictions(lg, th)
        assert np.array_equal(cls, golden["kp_cls" + tag]) and np.array_equal(idx, golden["kp_idx" + tag])
        assert np.allclose(pr.numpy(), golden["kp_pr" + tag], atol=2e-6)
    idx, cls, pr = O.get_key_point_predictions(torch.from_numpy(golden["kp10_logits"]), 0.75)
    assert np.array_equal(cls, golden["kp10_cls"]) and np.array_equal(idx, golden["kp10_idx"])
    vo = torch.from_numpy(golden["vote_out"])
    assert np.allclose(O.get_pred_center(vo, golden["vote_coords"], ee_r=0.02), golden["vote_center"], atol=1e-6)
    assert np.allclose(O.get_pred_center(vo, golden["vote_coords"], ee_r=0.02, q=golden["vote_q"]),
                       golden["vote_center_q"], atol=1e-6)
    # segmentation arg-max: the kernel the pipeline uses (K5 arg-max by-product) and the field-level helper
    import MinkowskiEngine as ME
    from MinkowskiEngine._lib import ptr, stream, check
    sl = torch.from_numpy(golden["seg_logits"]).cuda().contiguous()
    eye = torch.eye(3, device="cuda").contiguous()
    am = torch.empty((sl.shape[0],), dtype=torch.uint8, device="cuda")
    check(ME._C.b2me_linear_small(ptr(sl), 0, sl.shape[0], 3, ptr(eye), None, 3, None, ptr(am), stream()))
    assert np.array_equal(am.cpu().numpy().astype(np.int64), golden["seg_preds"])

    class _Field:
        features = sl
    preds, conf = O.get_segmentations_from_tensor_field(_Field())
    assert np.array_equal(preds, golden["seg_preds"]) and np.allclose(conf, golden["seg_conf"], atol=1e-6)


def test_translation_magic_vs_reference_golden(golden):
    """fused translation kernel against InferenceEngine.predict_translation of the reference itself; tolerance
    1e-4 m (north star), observed ~1e-6 (fp32 rotation like the reference, fp64 tail)."""
    from b200calib import output as O
    ns = golden["trans_n"]
    offs = np.concatenate(([0], np.cumsum(ns))).astype(np.int32)
    pts = torch.from_numpy(np.concatenate([golden["trans_pts"][i][:n] for i, n in enumerate(ns)])).cuda()
    out = O.translation_magic_batched(pts, offs, torch.from_numpy(golden["trans_q"]).cuda()).cpu().numpy()
    err = np.abs(out - golden["trans_out"]).max()
    print("translation max abs err vs reference", err)
    assert err < 1e-4
    assert err < 2e-5


def test_sanity_check_kernel_vs_reference_golden():
    """b2me_sanity_check (on-device InferenceEngine.check_sanity, app/inference_engine.py:246-279 + utils/data.py:255-335
    + utils/metrics.py:130-136) against the verdicts of the reference's OWN check_sanity on 24 seeded EE crops
    (tests/golden/make_golden_sanity.py: both verdicts, every failure mode), all crops in one launch."""
    import os
    from conftest import GOLDEN
    from b200calib.output import sanity_check_batched
    g = np.load(os.path.join(GOLDEN, "reference_sanity.npz"))
    S = len(g["sane"])
    crops, offs, prob, xyz = [], [0], np.zeros((S, 6), np.float32), np.zeros((S, 6, 3), np.float32)
    for c in range(S):
        ee = g["points"][c][g["seg"][c] == 2]
        crops.append(ee)
        offs.append(offs[-1] + len(ee))
        m = int(g["n_pred"][c])
        for k, p in zip(g["pred_cls"][c][:m], g["pred_xyz"][c][:m]):
            prob[c, int(k)] = 1.0
            xyz[c, int(k)] = p
    got = sanity_check_batched(torch.from_numpy(np.concatenate(crops)).cuda(), np.asarray(offs, np.int32),
                               torch.from_numpy(g["pose"]), torch.from_numpy(prob), torch.from_numpy(xyz),
                               kp_threshold=0.5).cpu().numpy().astype(bool)
    assert np.array_equal(got, g["sane"]), np.nonzero(got != g["sane"])[0]
    assert 5 < got.sum() < S - 5
    # without key points only the point-count and corner tests apply (len(result.key_points) <= 3)
    nokp = sanity_check_batched(torch.from_numpy(np.concatenate(crops)).cuda(), np.asarray(offs, np.int32),
                                torch.from_numpy(g["pose"])).cpu().numpy().astype(bool)
    from b200calib import sanity as HS
    want = np.array([HS.check_sanity(g["points"][c], g["seg"][c].astype(np.int64), g["pose"][c], []) for c in range(S)])
    assert np.array_equal(nokp, want)


def test_normalize_colors_kernel_vs_reference_golden(golden):
    """b2me_normalize_colors (utils/preprocess.py:20-37 per frame of a batch) against outputs of the reference's own
    normalize_colors: 0..255 input, 0..1 input, negative inputs (per-channel min-max branch) and an already centred
    input as FIVE FRAMES OF ONE BATCH (mixed ranges in one batch), bit-exact."""
    from b200calib.output import normalize_colors_batched
    keys = ["pre_rgb255", "pre_rgb01", "pre_rgbneg", "pre_rgbcen", "pre_rgbneg255"]
    frames = [golden[k] for k in keys]
    offs = np.concatenate(([0], np.cumsum([len(f) for f in frames]))).astype(np.int32)
    bidx = np.repeat(np.arange(len(frames), dtype=np.float32), [len(f) for f in frames])
    out = normalize_colors_batched(torch.from_numpy(np.concatenate(frames)).cuda(), torch.from_numpy(bidx).cuda(),
                                   offs).cpu().numpy()
    for i, k in enumerate(keys):
        assert np.array_equal(out[offs[i]:offs[i + 1]], golden[k + "_out"]), k
