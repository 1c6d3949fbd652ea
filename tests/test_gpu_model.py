"""Model-level parity: the same MinkUNet18D-based networks on the CUDA package vs the CPU oracle package,
same state dict, same synthetic frame. fp32 path: features within 1e-3 relative, labels identical wherever the
oracle's top-2 logit margin exceeds the numerical noise; bf16 (tcgen05) path: features within 2e-2."""
import numpy as np
import pytest
import torch

import oracle.MinkowskiEngine as OME
from b200calib.models import make_models, randomize_bn_stats
from b200calib.synthetic import make_frame
from gpu_util import rel_err

pytestmark = pytest.mark.gpu


def _frame(width=160, height=120, seed=13):
    f = make_frame(seed, width=width, height=height)
    return torch.from_numpy(f["points"]), torch.from_numpy(f["rgb"]) - 0.5


def _run(ME, net, pts, rgb, scale, device=None):
    co = OME.utils.batched_coordinates([p * scale for p in pts], dtype=torch.float32)
    fe = torch.cat(rgb)
    kw = dict(device=device) if device else {}
    fld = ME.TensorField(features=fe, coordinates=co, quantization_mode=ME.SparseTensorQuantizationMode.UNWEIGHTED_AVERAGE,
                         minkowski_algorithm=ME.MinkowskiAlgorithm.SPEED_OPTIMIZED, **kw)
    with torch.no_grad():
        out = net(fld.sparse())
    return out, fld


@pytest.fixture(scope="module")
def nets():
    import MinkowskiEngine as ME
    torch.manual_seed(13)
    MO, MC = make_models(OME), make_models(ME)
    o = randomize_bn_stats(MO.RobotNetSegmentation(3, num_classes=3)).eval()
    with torch.no_grad():   # spread the three classes so that the arg-max is not constant
        o.regression[2].linear.bias.copy_(torch.tensor([0.0, 0.02, -0.02]))
    c = MC.RobotNetSegmentation(3, num_classes=3)
    c.load_state_dict(o.state_dict())
    return ME, o, c.cuda().eval()


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-3), ("tf32", 1e-3), (torch.bfloat16, 2e-2)])
def test_segmentation_forward_parity(nets, dtype, tol):
    ME, onet, cnet = nets
    pts, rgb = zip(_frame(seed=13), _frame(seed=14))
    ME.set_compute_dtype(dtype)
    try:
        oo, of = _run(OME, onet, pts, rgb, 100.0)
        co, cf = _run(ME, cnet, pts, rgb, 100.0, device="cuda")
        assert torch.equal(co.C.cpu(), oo.C)
        lo, lc = oo.F, co.F.float().cpu()
        err = rel_err(lc, lo)
        print(f"logits rel err ({dtype}): {err:.3e}")
        assert err < tol
        po, pc = oo.slice(of).F, co.slice(cf).F.float().cpu()
        lab_o, lab_c = po.max(1)[1], pc.max(1)[1]
        top2 = po.topk(2, dim=1)[0]
        margin = top2[:, 0] - top2[:, 1]
        noise = (pc - po).abs().max()
        decided = margin > 2 * noise
        assert torch.equal(lab_o[decided], lab_c[decided])
        mism = int((lab_o != lab_c).sum())
        print(f"labels ({dtype}): {mism} of {len(lab_o)} differ from the oracle (no mask)")
        if dtype == torch.float32:
            assert float(decided.float().mean()) > 0.99
            assert float((lab_o == lab_c).float().mean()) > 0.999
    finally:
        ME.set_compute_dtype(torch.float32)


def test_backbone_features_parity_fp32(nets):
    """UNet trunk output (256-channel features before the head), fp32: <= 1e-3 relative."""
    ME, onet, cnet = nets
    pts, rgb = _frame(seed=15)
    co_ = OME.utils.batched_coordinates([pts * 100.0], dtype=torch.float32)
    with torch.no_grad():
        o = onet.final(onet.forward_except_final(OME.TensorField(features=rgb, coordinates=co_).sparse()))
        c = cnet.final(cnet.forward_except_final(ME.TensorField(features=rgb, coordinates=co_, device="cuda").sparse()))
    assert rel_err(c.F, o.F) < 1e-3


def test_encode_parity(nets):
    ME, _, _ = nets
    torch.manual_seed(5)
    MO, MC = make_models(OME), make_models(ME)
    o = randomize_bn_stats(MO.RobotNetEncode(3, 7)).eval()
    c = MC.RobotNetEncode(3, 7)
    c.load_state_dict(o.state_dict())
    c = c.cuda().eval()
    pts = [_frame(seed=s)[0][:3000] * 0.2 for s in (1, 2, 3)]
    rgb = [_frame(seed=s)[1][:3000] for s in (1, 2, 3)]
    oo, _ = _run(OME, o, pts, rgb, 200.0)
    co, _ = _run(ME, c, pts, rgb, 200.0, device="cuda")
    assert oo.shape == (3, 7)
    assert torch.allclose(co.cpu(), oo, atol=2e-4)
    ME.set_compute_dtype(torch.bfloat16)
    try:
        cb, _ = _run(ME, c, pts, rgb, 200.0, device="cuda")
        assert torch.allclose(cb.cpu(), oo, atol=3e-2)
    finally:
        ME.set_compute_dtype(torch.float32)


def test_lazy_fusion_launch_count(nets):
    """conv+BN+ReLU(+residual) chains and ME.cat must fuse: one launch per convolution, none per BN/ReLU/cat."""
    ME, _, cnet = nets
    pts, rgb = _frame(seed=16)
    co_ = OME.utils.batched_coordinates([pts * 100.0], dtype=torch.float32)
    x = ME.TensorField(features=rgb, coordinates=co_, device="cuda").sparse()
    with torch.no_grad():
        cnet(x).F                    # builds maps and weight caches
        x2 = ME.SparseTensor(features=x.F, coordinate_map_key=x.coordinate_map_key,
                             coordinate_manager=x.coordinate_manager)
        ME.reset_launch_count()
        cnet(x2).F
    # 49 convolutions + 2 head linears (maps cached): no separate BN / ReLU / add / cat kernels
    assert ME.launch_count() == 51
    # tensor-core modes: the two head linears are ONE launch (b2me_head_fused_tc); + 1 conversion of the fp32 voxel
    # features of the test's SparseTensor to the activation type... none: the stem runs on the SIMT kernel from fp32
    for mode in ("bf16", "tf32"):
        ME.set_compute_dtype(mode)
        try:
            with torch.no_grad():
                cnet(x2).F           # packs the weights of this operand type
                ME.reset_launch_count()
                out = cnet(x2)
                out.F
            assert ME.launch_count() == 50, (mode, ME.launch_count())
            assert getattr(out, "_row_argmax", None) is not None
            ME.set_fuse_head(False)
            with torch.no_grad():
                ME.reset_launch_count()
                two = cnet(x2)
                two.F
            assert ME.launch_count() == 51
            assert float((two.F - out.F).abs().max()) <= 1e-5 * float(two.F.abs().max())
            assert int((two._row_argmax != out._row_argmax).sum()) <= 1
        finally:
            ME.set_fuse_head(True)
            ME.set_compute_dtype(torch.float32)


def test_operand_paths_bit_identical_at_model_level(nets):
    """the whole segmentation forward through the TMA operand path (default) and the cp.async path: same bits, in both
    tensor-core modes (each output row is one fp32 accumulation in a fixed (offset, chunk) order on either path)."""
    ME, _, cnet = nets
    pts, rgb = _frame(seed=21)
    co_ = OME.utils.batched_coordinates([pts * 100.0], dtype=torch.float32)
    for mode in ("bf16", "tf32"):
        ME.set_compute_dtype(mode)
        try:
            outs = {}
            for path in ("tma", "cpasync"):
                ME.set_tc_operand_path(path)
                assert ME.get_tc_operand_path() == path
                x = ME.TensorField(features=rgb, coordinates=co_, device="cuda").sparse()
                with torch.no_grad():
                    outs[path] = cnet(x).F.clone()
            assert torch.equal(outs["tma"], outs["cpasync"]), (mode, float((outs["tma"] - outs["cpasync"]).abs().max()))
        finally:
            ME.set_tc_operand_path("tma")
            ME.set_compute_dtype(torch.float32)


def test_robotnet_full_unet_global_max_pool_parity(nets):
    """model/robotnet.py (full UNet -> BN+ReLU -> MinkowskiGlobalMaxPooling -> MLP)."""
    ME, _, _ = nets
    torch.manual_seed(7)
    MO, MC = make_models(OME), make_models(ME)
    o = randomize_bn_stats(MO.RobotNet(3, 7)).eval()
    c = MC.RobotNet(3, 7)
    c.load_state_dict(o.state_dict())
    c = c.cuda().eval()
    pts = [_frame(seed=s)[0][:4000] * 0.25 for s in (4, 5)]
    rgb = [_frame(seed=s)[1][:4000] for s in (4, 5)]
    oo, _ = _run(OME, o, pts, rgb, 200.0)
    co, _ = _run(ME, c, pts, rgb, 200.0, device="cuda")
    assert oo.shape == (2, 7)
    assert torch.allclose(co.cpu(), oo, atol=5e-4), float((co.cpu() - oo).abs().max())
    ME.set_compute_dtype(torch.bfloat16)
    try:
        cb, _ = _run(ME, c, pts, rgb, 200.0, device="cuda")
        assert torch.allclose(cb.cpu(), oo, atol=5e-2), float((cb.cpu() - oo).abs().max())
    finally:
        ME.set_compute_dtype(torch.float32)


@pytest.mark.parametrize("variant", ["MinkUNet50", "MinkUNet14A", "MinkUNet34C"])
def test_other_minkunet_trunks_parity(nets, variant):
    """Bottleneck trunks (expansion 4: 1x1 convolutions dominate, channel counts 128..1024)
    and the other BasicBlock variants robotnet_segmentation.py:17-28 can select, on the same kernels."""
    ME, _, _ = nets
    torch.manual_seed(11)
    MO, MC = make_models(OME), make_models(ME)
    o = randomize_bn_stats(MO.RobotNetSegmentation(3, num_classes=3, variant=variant)).eval()
    c = MC.RobotNetSegmentation(3, num_classes=3, variant=variant)
    c.load_state_dict(o.state_dict())
    c = c.cuda().eval()
    pts, rgb = zip(_frame(width=96, height=72, seed=21))
    oo, _ = _run(OME, o, pts, rgb, 100.0)
    for dtype, tol in ((torch.float32, 1e-3), (torch.bfloat16, 3e-2)):
        ME.set_compute_dtype(dtype)
        try:
            co, _ = _run(ME, c, pts, rgb, 100.0, device="cuda")
            assert torch.equal(co.C.cpu(), oo.C)
            err = rel_err(co.F.float().cpu(), oo.F)
            print(f"{variant} {dtype}: logits rel err {err:.3e}")
            assert err < tol, (variant, dtype, err)
        finally:
            ME.set_compute_dtype(torch.float32)


def test_aliveunet_trunk_parity(nets):
    """AliveUNet (model/backbone/aliveunet.py:45-275, the fall-back trunk of the robotnet
    models): seven stride-2 levels down to tensor stride 128 and back, channel counts 32..416 in steps of 32, on the
    same kernels. The mirror equals the unchanged reference class on the oracle (tests/ref_model_runner.py)."""
    ME, _, _ = nets
    torch.manual_seed(17)
    MO, MC = make_models(OME), make_models(ME)
    o = randomize_bn_stats(MO.AliveUNet(3, 7, m=32, block_reps=1)).eval()
    c = MC.AliveUNet(3, 7, m=32, block_reps=1)
    c.load_state_dict(o.state_dict())
    c = c.cuda().eval()
    pts, rgb = zip(_frame(width=128, height=96, seed=23))
    oo, _ = _run(OME, o, pts, rgb, 100.0)
    for dtype, tol in ((torch.float32, 1e-3), (torch.bfloat16, 3e-2)):
        ME.set_compute_dtype(dtype)
        try:
            co, _ = _run(ME, c, pts, rgb, 100.0, device="cuda")
            assert torch.equal(co.C.cpu(), oo.C)
            err = rel_err(co.F.float().cpu(), oo.F)
            print(f"AliveUNet {dtype}: features rel err {err:.3e}")
            assert err < tol, (dtype, err)
        finally:
            ME.set_compute_dtype(torch.float32)
