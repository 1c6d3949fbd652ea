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


def _call_tc(lib, x, x2, W, nbr, V_out, scale, shift, res, act, out_dtype=torch.float32, perm=None):
    from MinkowskiEngine._lib import ptr, stream, dtype_code, check
    K, _, Cout = W.shape
    c1, c2 = x.shape[1], 0 if x2 is None else x2.shape[1]
    assert lib.b2me_tc_supported(K, c1, c2, Cout) == 1
    packed = torch.empty((lib.b2me_tc_packed_bytes(K, c1, c2, Cout),), dtype=torch.uint8, device="cuda")
    check(lib.b2me_tc_pack_weights(ptr(W), K, c1, c2, Cout, ptr(packed), stream()))
    out = torch.empty((V_out, Cout), dtype=out_dtype, device="cuda")
    import MinkowskiEngine as ME
    masks = ME.tile_masks(nbr, perm, V_out, K) if nbr is not None else None
    check(lib.b2me_spconv_fwd_tc(ptr(x), c1, ptr(x2), c2, x.shape[0], ptr(packed), ptr(nbr), ptr(perm), ptr(masks), K, V_out, Cout,
                                 ptr(scale), ptr(shift), ptr(res), act, 0.01, ptr(out), dtype_code(out_dtype),
                                 stream()))
    torch.cuda.synchronize()
    return out


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


@pytest.mark.parametrize("kind,c1,c2,cout,epi", CASES)
def test_spconv_parity(kind, c1, c2, cout, epi):
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

    # ---- SIMT fp32: tolerance 1e-3 relative (north star), observed ~1e-6
    ref = _oracle(xin, W, nbr_o, V_out, scale, shift, res, act)
    got = _call_simt(lib, cu(x), cu(x2), cu(W), nbr_c, V_out, cu(scale), cu(shift), cu(res), act)
    assert rel_err(got, ref) < 1e-5
    assert float((got.cpu() - ref).abs().max()) < 1e-3 * float(ref.abs().max())

    # ---- tcgen05: operands rounded to bf16 exactly as the kernel sees them, fp32 accumulation
    if lib.b2me_tc_supported(K, c1, c2, cout):
        b16 = lambda t: None if t is None else t.bfloat16()
        refb = _oracle(b16(xin).float(), b16(W).float(), nbr_o, V_out, scale, shift,
                       None if res is None else b16(res).float(), act)
        gotb = _call_tc(lib, cu(b16(x)), cu(b16(x2)), cu(W), nbr_c, V_out, cu(scale), cu(shift), cu(b16(res)), act)
        assert rel_err(gotb, refb) < 1e-5, "tcgen05 path (fp32 out) differs from bf16-operand oracle"
        gotb16 = _call_tc(lib, cu(b16(x)), cu(b16(x2)), cu(W), nbr_c, V_out, cu(scale), cu(shift), cu(b16(res)), act,
                          out_dtype=torch.bfloat16)
        assert rel_err(gotb16.float(), refb) < 4e-3          # one bf16 rounding of the output
        assert rel_err(gotb16.float(), ref) < 2e-2           # north-star bf16 tolerance vs the fp32 oracle
    else:
        assert c1 < 16 or cout % 16


@pytest.mark.parametrize("kind,c1,cout", [("k3", 64, 128), ("up", 128, 64), ("down", 32, 32)])
def test_tc_row_permutation_is_bit_identical(kind, c1, cout):
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
    base = _call_tc(lib, x, None, W, nbr_c, V_out, sc, None, res, 1)
    perm = ME.mask_sorted_perm(nbr_c, V_out, K)
    assert torch.equal(torch.sort(perm.long())[0].cpu(), torch.arange(V_out))
    # keys are non-decreasing along perm: rows with equal masks are adjacent
    got = _call_tc(lib, x, None, W, nbr_c, V_out, sc, None, res, 1, perm=perm)
    assert torch.equal(got, base)
    rnd = torch.randperm(V_out, generator=g).int().cuda()
    assert torch.equal(_call_tc(lib, x, None, W, nbr_c, V_out, sc, None, res, 1, perm=rnd), base)
    # the sorted order needs fewer non-empty (tile, offset) pairs than the natural one
    pres = (nbr_c >= 0).cpu()

    def tiles(order):
        pad = (-V_out) % 128
        p_ = torch.cat((pres[order], torch.zeros(pad, K, dtype=torch.bool)))
        return int(p_.view(-1, 128, K).any(1).sum())
    assert tiles(perm.long().cpu()) < tiles(torch.arange(V_out))


def test_tc_small_and_ragged_tiles():
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
        got = _call_tc(lib, x.cuda(), None, W.cuda(), nbr.cuda(), V, None, None, None, 0)
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
