"""K4 parity: SIMT (fp32) and tcgen05 (bf16 operands, fp32 accumulation) sparse convolution against the
oracle's gather-matmul-index_add, through the C ABI, on real kernel maps."""
import numpy as np
import pytest
import torch

import oracle.MinkowskiEngine as OME
from gpu_util import kinect_like_cloud, rel_err

pytestmark = pytest.mark.gpu


def _setup(n=30000, scale=60.0, seed=5):
    import MinkowskiEngine as ME
    pts = kinect_like_cloud(n, seed)
    co = OME.utils.batched_coordinates([torch.from_numpy(pts) * scale], dtype=torch.float32)
    fe = torch.zeros(len(pts), 1)
    os_ = OME.TensorField(features=fe, coordinates=co).sparse()
    cs = ME.TensorField(features=fe, coordinates=co, device="cuda").sparse()
    return ME, os_, cs


def _maps(os_, cs, kind):
    om, cm = os_.coordinate_manager, cs.coordinate_manager
    ok, ck = os_.coordinate_map_key, cs.coordinate_map_key
    if kind == "k3":
        return om.kernel_map_k3(ok), cm.kernel_map_k3(ck), om.levels[ok].V, om.levels[ok].V
    ok2, ro = om.stride_down(ok)
    ck2, rc = cm.stride_down(ck)
    if kind == "down":
        return ro["nbr_down"], rc["nbr_down"], om.levels[ok].V, om.levels[ok2].V
    if kind == "up":
        return ro["nbr_up"], rc["nbr_up"], om.levels[ok2].V, om.levels[ok].V
    return None, None, om.levels[ok].V, om.levels[ok].V   # identity (k1)


def _call_simt(lib, x, x2, W, nbr, V_out, scale, shift, res, act, out_dtype=torch.float32):
    from MinkowskiEngine._lib import ptr, stream, dtype_code, check
    K, _, Cout = W.shape
    out = torch.empty((V_out, Cout), dtype=out_dtype, device="cuda")
    check(lib.b2me_spconv_fwd_simt(ptr(x), x.shape[1], ptr(x2), 0 if x2 is None else x2.shape[1], dtype_code(x.dtype),
                                   ptr(W), ptr(nbr), K, V_out, Cout, ptr(scale), ptr(shift), ptr(res),
                                   0 if res is None else dtype_code(res.dtype), act, 0.01, ptr(out),
                                   dtype_code(out_dtype), stream()))
    return out


def _call_tc(lib, x, x2, W, nbr, V_out, scale, shift, res, act, out_dtype=torch.float32, perm=None, path="cpasync",
             tf32=False, rot128=True, round_out=True):
    """b2me_spconv_fwd_tc through ctypes. bf16 operands (default; x / res bf16) or tf32 operands (x / res fp32)."""
    from MinkowskiEngine import _lib
    from MinkowskiEngine._lib import ptr, stream, check
    K, _, Cout = W.shape
    c1, c2 = x.shape[1], 0 if x2 is None else x2.shape[1]
    assert lib.b2me_tc_supported(K, c1, c2, Cout) == 1
    op = _lib.TF32 if tf32 else _lib.BF16
    packed = torch.empty((lib.b2me_tc_packed_bytes(K, c1, c2, Cout, op),), dtype=torch.uint8, device="cuda")
    check(lib.b2me_tc_pack_weights(ptr(W), K, c1, c2, Cout, op, ptr(packed), stream()))
    out = torch.empty((V_out, Cout), dtype=out_dtype, device="cuda")
    import MinkowskiEngine as ME
    masks = ME.tile_masks(nbr, perm, V_out, K) if nbr is not None else None
    flags = (_lib.TC_FLAG_TMA if path == "tma" else 0) | (0 if rot128 else _lib.TC_FLAG_NO_ROT128)
    out_code = _lib.BF16 if out_dtype == torch.bfloat16 else (_lib.TF32 if (tf32 and round_out) else _lib.F32)
    check(lib.b2me_spconv_fwd_tc(ptr(x), c1, ptr(x2), c2, x.shape[0], op, ptr(packed), ptr(nbr), ptr(perm), ptr(masks),
                                 K, V_out, Cout, ptr(scale), ptr(shift), ptr(res), act, 0.01, ptr(out), out_code,
                                 flags, stream()))
    torch.cuda.synchronize()
    return out


def _tf32(t):
    """round to nearest tf32 (10-bit mantissa, ties away from zero like cvt.rna.tf32.f32), fp32 container."""
    if t is None:
        return None
    u = t.contiguous().view(torch.int32).to(torch.int64) & 0xFFFFFFFF
    u = ((u + 0x1000) & 0xFFFFE000) & 0xFFFFFFFF
    u = torch.where(u >= 2 ** 31, u - 2 ** 32, u).to(torch.int32)
    return u.view(torch.float32).reshape(t.shape)


def _oracle(x, W, nbr, V_out, scale, shift, res, act):
    out = OME.sparse_conv(x, W, nbr, V_out)
    if scale is not None:
        out = out * scale
    if shift is not None:
        out = out + shift
    if res is not None:
        out = out + res
    if act == 1:
        out = torch.relu(out)
    elif act == 2:
        out = torch.nn.functional.leaky_relu(out, 0.01)
    return out


CASES = [  # kind, Cin1, Cin2, Cout, epilogue
    ("k3", 3, 0, 32, "bn_relu"),
    ("k3", 32, 0, 32, "none"),
    ("k3", 64, 0, 128, "bn_res_relu"),
    ("down", 32, 0, 32, "bn_relu"),
    ("up", 256, 0, 384, "bn_relu"),
    ("k3", 384, 32, 384, "bn_relu"),       # ME.cat consumed as two sources (416 -> 384)
    ("k1", 384, 64, 384, "bn"),            # downsample 1x1 on a cat
    ("k1", 384, 0, 256, "bias_leaky"),     # final conv + leaky relu
    ("k1", 256, 0, 1024, "bias_leaky"),    # head Linear 256 -> 1024
    ("k3", 128, 0, 256, "bn_res_relu"),
    ("k3", 48, 16, 80, "bn_relu"),         # partial chunks (kw = 48 / 16), odd N
]


PATHS = ["cpasync", "tma"]   # the two operand paths of k_spconv_tc (ME.set_tc_operand_path / B2ME_TC_FLAG_TMA)


@pytest.mark.parametrize("path", PATHS)
@pytest.mark.parametrize("kind,c1,c2,cout,epi", CASES)
def test_spconv_parity(kind, c1, c2, cout, epi, path):
    ME, os_, cs = _setup()
    lib = ME._C
    nbr_o, nbr_c, V_in, V_out = _maps(os_, cs, kind)
    g = torch.Generator().manual_seed(c1 * 7 + cout)
    K = {"k3": 27, "down": 8, "up": 8, "k1": 1}[kind]
    x = torch.randn(V_in, c1, generator=g)
    x2 = torch.randn(V_in, c2, generator=g) if c2 else None
    W = torch.randn(K, c1 + c2, cout, generator=g) / np.sqrt(K * (c1 + c2) / 3)
    scale = torch.rand(cout, generator=g) + 0.5 if "bn" in epi else None
    shift = torch.randn(cout, generator=g) * 0.1 if ("bn" in epi or "bias" in epi) else None
    res = torch.randn(V_out, cout, generator=g) if "res" in epi else None
    act = 1 if "relu" in epi else (2 if "leaky" in epi else 0)
    xin = torch.cat((x, x2), 1) if c2 else x
    cu = lambda t: None if t is None else t.cuda().contiguous()

    ref = _oracle(xin, W, nbr_o, V_out, scale, shift, res, act)
    if path == "cpasync":
        # ---- SIMT fp32: tolerance 1e-3 relative (north star), observed ~1e-6
        got = _call_simt(lib, cu(x), cu(x2), cu(W), nbr_c, V_out, cu(scale), cu(shift), cu(res), act)
        assert rel_err(got, ref) < 1e-5
        assert float((got.cpu() - ref).abs().max()) < 1e-3 * float(ref.abs().max())

    if not lib.b2me_tc_supported(K, c1, c2, cout):
        assert c1 < 16 or cout % 16
        return
    # ---- tcgen05, bf16 operands rounded exactly as the kernel sees them, fp32 accumulation
    b16 = lambda t: None if t is None else t.bfloat16()
    refb = _oracle(b16(xin).float(), b16(W).float(), nbr_o, V_out, scale, shift,
                   None if res is None else b16(res).float(), act)
    gotb = _call_tc(lib, cu(b16(x)), cu(b16(x2)), cu(W), nbr_c, V_out, cu(scale), cu(shift), cu(b16(res)), act,
                    path=path)
    assert rel_err(gotb, refb) < 1e-5, "tcgen05 path (fp32 out) differs from bf16-operand oracle"
    gotb16 = _call_tc(lib, cu(b16(x)), cu(b16(x2)), cu(W), nbr_c, V_out, cu(scale), cu(shift), cu(b16(res)), act,
                      out_dtype=torch.bfloat16, path=path)
    assert rel_err(gotb16.float(), refb) < 4e-3          # one bf16 rounding of the output
    assert rel_err(gotb16.float(), ref) < 2e-2           # north-star bf16 tolerance vs the fp32 oracle
    if cout == 384 and K != 1:
        # 384-column tiles: the early-release / alternating-region accumulator layout changes no bit
        single = _call_tc(lib, cu(b16(x)), cu(b16(x2)), cu(W), nbr_c, V_out, cu(scale), cu(shift), cu(b16(res)), act,
                          out_dtype=torch.bfloat16, path=path, rot128=False)
        assert torch.equal(single, gotb16)

    # ---- tcgen05 kind::tf32: fp32 rows holding tf32 values, fp32 accumulation; fp32 tolerance of the north star
    reft = _oracle(_tf32(xin), _tf32(W), nbr_o, V_out, scale, shift, res, act)
    gott = _call_tc(lib, cu(_tf32(x)), cu(_tf32(x2)), cu(W), nbr_c, V_out, cu(scale), cu(shift), cu(res), act,
                    path=path, tf32=True, round_out=False)
    assert rel_err(gott, reft) < 1e-5, "tf32 tensor-core path differs from the tf32-operand oracle"
    gotr = _call_tc(lib, cu(_tf32(x)), cu(_tf32(x2)), cu(W), nbr_c, V_out, cu(scale), cu(shift), cu(res), act,
                    path=path, tf32=True)
    assert torch.equal(gotr.cpu(), _tf32(gott.cpu())), "B2ME_TF32 output = the fp32 result rounded to nearest tf32"
    assert rel_err(gotr, ref) < 1e-3                     # north-star fp32 tolerance vs the fp32 oracle


@pytest.mark.parametrize("path", PATHS)
@pytest.mark.parametrize("kind,c1,cout", [("k3", 64, 128), ("up", 128, 64), ("down", 32, 32), ("k3", 96, 384)])
def test_tc_row_permutation_is_bit_identical(kind, c1, cout, path):
    """K3b: any row permutation (mask-sorted or random) must give bit-identical outputs - absent neighbours
    contribute exact zeros - while the mask-sorted one skips (tile, offset) passes."""
    ME, os_, cs = _setup(n=20000)
    lib = ME._C
    nbr_o, nbr_c, V_in, V_out = _maps(os_, cs, kind)
    K = nbr_c.shape[1]
    g = torch.Generator().manual_seed(11)
    x = torch.randn(V_in, c1, generator=g).bfloat16().cuda()
    W = (torch.randn(K, c1, cout, generator=g) / 8).cuda()
    res = torch.randn(V_out, cout, generator=g).bfloat16().cuda()
    sc = (torch.rand(cout, generator=g) + 0.5).cuda()
    base = _call_tc(lib, x, None, W, nbr_c, V_out, sc, None, res, 1, path=path)
    # the two operand paths agree bit for bit
    other = _call_tc(lib, x, None, W, nbr_c, V_out, sc, None, res, 1, path=[q for q in PATHS if q != path][0])
    assert torch.equal(other, base)
    perm = ME.mask_sorted_perm(nbr_c, V_out, K)
    assert torch.equal(torch.sort(perm.long())[0].cpu(), torch.arange(V_out))
    # keys are non-decreasing along perm: rows with equal masks are adjacent
    got = _call_tc(lib, x, None, W, nbr_c, V_out, sc, None, res, 1, perm=perm, path=path)
    assert torch.equal(got, base)
    rnd = torch.randperm(V_out, generator=g).int().cuda()
    assert torch.equal(_call_tc(lib, x, None, W, nbr_c, V_out, sc, None, res, 1, perm=rnd, path=path), base)
    gotb = _call_tc(lib, x, None, W, nbr_c, V_out, sc, None, res, 1, perm=perm, path=path, out_dtype=torch.bfloat16)
    assert torch.equal(gotb, _call_tc(lib, x, None, W, nbr_c, V_out, sc, None, res, 1, perm=rnd, path=path,
                                      out_dtype=torch.bfloat16))
    # the sorted order needs fewer non-empty (tile, offset) pairs than the natural one
    pres = (nbr_c >= 0).cpu()

    def tiles(order):
        pad = (-V_out) % 128
        p_ = torch.cat((pres[order], torch.zeros(pad, K, dtype=torch.bool)))
        return int(p_.view(-1, 128, K).any(1).sum())
    assert tiles(perm.long().cpu()) < tiles(torch.arange(V_out))


@pytest.mark.parametrize("path", PATHS)
def test_tc_small_and_ragged_tiles(path):
    """V_out not a multiple of 128, tiny V, and a tile whose rows have no neighbours except themselves."""
    import MinkowskiEngine as ME
    lib = ME._C
    g = torch.Generator().manual_seed(1)
    for V in (1, 127, 129, 1000):
        nbr = torch.full((V, 27), -1, dtype=torch.int32)
        nbr[:, 13] = torch.arange(V, dtype=torch.int32)
        if V > 2:
            nbr[1:, 12] = torch.arange(V - 1, dtype=torch.int32)
        x = torch.randn(V, 64, generator=g).bfloat16()
        W = torch.randn(27, 64, 64, generator=g) / 8
        ref = OME.sparse_conv(x.float(), W.bfloat16().float(), nbr.numpy().astype(np.int64), V)
        got = _call_tc(lib, x.cuda(), None, W.cuda(), nbr.cuda(), V, None, None, None, 0, path=path)
        assert rel_err(got, ref) < 1e-5


def test_linear_small_and_gathers():
    import MinkowskiEngine as ME
    from MinkowskiEngine._lib import ptr, stream, check
    lib = ME._C
    g = torch.Generator().manual_seed(2)
    V = 5000
    x = torch.randn(V, 1024, generator=g)
    lin = torch.nn.Linear(1024, 3)
    ref = lin(x).detach()
    Wc, bc = lin.weight.detach().cuda().contiguous(), lin.bias.detach().cuda()   # keep alive across the launch
    for xin in (x, x.bfloat16()):
        out = torch.empty((V, 3), device="cuda")
        am = torch.empty((V,), dtype=torch.uint8, device="cuda")
        xc = xin.cuda()
        check(lib.b2me_linear_small(ptr(xc), 0 if xin.dtype == torch.float32 else 1, V, 1024,
                                    ptr(Wc), ptr(bc), 3,
                                    ptr(out), ptr(am), stream()))
        r = lin(xin.float()).detach()
        assert rel_err(out, r) < 1e-5
        assert torch.equal(am.cpu().long(), out.cpu().max(1)[1])
    assert rel_err(out, ref) < 2e-2
    inv = torch.randint(0, V, (20000,), generator=g).int()
    lab = torch.randint(0, 3, (V,), generator=g).to(torch.uint8)
    outl = torch.empty((20000,), dtype=torch.uint8, device="cuda")
    labc, invc = lab.cuda(), inv.cuda()
    check(lib.b2me_gather_labels(ptr(labc), ptr(invc), 20000, ptr(outl), stream()))
    assert torch.equal(outl.cpu(), lab[inv.long()])


@pytest.mark.parametrize("path", PATHS)
@pytest.mark.parametrize("V,C2,mode", [(70000, 3, "bf16"), (5000, 6, "bf16"), (257, 10, "bf16"), (70000, 3, "tf32"),
                                       (1000, 16, "tf32")])
def test_fused_head_matches_two_launch_path(V, C2, mode, path):
    """b2me_head_fused_tc (256 -> 1024 -> C2 in one launch, hidden activation on chip) against the two-launch data
    path it replaces (K = 1 tcgen05 linear storing the hidden tensor, then b2me_linear_small) and the fp32 oracle
    (model/robotnet_segmentation.py:43-49). The only difference allowed between the two CUDA paths is the fp32
    summation order of the 1024-term dot products; labels are compared without any margin mask."""
    import MinkowskiEngine as ME
    from MinkowskiEngine import _lib
    from MinkowskiEngine._lib import ptr, stream, check
    lib = ME._C
    g = torch.Generator().manual_seed(V + C2)
    tf32 = mode == "tf32"
    x = torch.randn(V, 256, generator=g)
    lin1, lin2 = torch.nn.Linear(256, 1024), torch.nn.Linear(1024, C2)
    with torch.no_grad():
        ref = lin2(torch.nn.functional.leaky_relu(lin1(x), 0.01))
    xq = _tf32(x) if tf32 else x.bfloat16()
    xc = xq.cuda().contiguous()
    W1 = lin1.weight.detach().t().contiguous().unsqueeze(0).cuda()       # [1, 256, 1024]
    b1 = lin1.bias.detach().cuda().contiguous()
    # two launches: hidden tensor in HBM
    hid = _call_tc(lib, xc, None, W1, None, V, None, b1, None, 2, out_dtype=torch.float32 if tf32 else torch.bfloat16,
                   path=path, tf32=tf32)
    W2t, b2 = lin2.weight.detach().cuda().contiguous(), lin2.bias.detach().cuda().contiguous()
    two = torch.empty((V, C2), device="cuda")
    two_am = torch.empty((V,), dtype=torch.uint8, device="cuda")
    check(lib.b2me_linear_small(ptr(hid), 0 if tf32 else 1, V, 1024, ptr(W2t), ptr(b2), C2, ptr(two), ptr(two_am),
                                stream()))
    # one launch
    op = _lib.TF32 if tf32 else _lib.BF16
    packed = torch.empty((lib.b2me_tc_packed_bytes(1, 256, 0, 1024, op),), dtype=torch.uint8, device="cuda")
    check(lib.b2me_tc_pack_weights(ptr(W1), 1, 256, 0, 1024, op, ptr(packed), stream()))
    C2p = (C2 + 3) // 4 * 4
    W2p = torch.zeros((1024, C2p), device="cuda")
    W2p[:, :C2] = W2t.t()
    one = torch.empty((V, C2), device="cuda")
    one_am = torch.empty((V,), dtype=torch.uint8, device="cuda")
    check(lib.b2me_head_fused_tc(ptr(xc), 256, V, op, ptr(packed), 1024, None, ptr(b1), 2, 0.01, ptr(W2p), ptr(b2), C2,
                                 ptr(one), ptr(one_am), _lib.TC_FLAG_TMA if path == "tma" else 0, stream()))
    torch.cuda.synchronize()
    assert rel_err(one, two) < 2e-6, "fused head differs from the two-launch path by more than fp32 summation order"
    assert torch.equal(one_am.cpu().long(), one.cpu().max(1)[1]), "arg-max is the lowest index of the row maximum"
    mism = int((one_am != two_am).sum())
    print(f"fused vs two-launch head ({mode}, {path}): {mism} of {V} labels differ, logits rel err "
          f"{rel_err(one, two):.2e}; vs fp32 oracle {rel_err(one, ref):.2e}")
    assert mism <= max(1, V // 20000), "labels of the fused head differ from the two-launch path"
    assert rel_err(one, ref) < (1e-3 if tf32 else 2e-2)


def test_tc_operand_path_switch_is_an_api():
    """ME.set_tc_operand_path selects the operand path of every following convolution (no environment variable)."""
    import MinkowskiEngine as ME
    assert ME.get_tc_operand_path() == "tma"       # the default (faster in the round-2 A/B, profiles/r02_ab_medians.md)
    ME.set_tc_operand_path("cpasync")
    assert ME.get_tc_operand_path() == "cpasync"
    ME.set_tc_operand_path("tma")
    with pytest.raises(ValueError):
        ME.set_tc_operand_path("ldg")
