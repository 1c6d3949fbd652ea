"""Execution of pending fused ops through libb2me (K4 + epilogues, K5, K6)."""
import torch

from . import _lib
from ._lib import lib, check, ptr, stream, dtype_code
from .core import SparseTensor, _Pending, _State, _as_dtype, _count, CoordinateMapKey, get_profile


# ------------------------------------------------------------------------------------------------ caches
def _tensor_version(t):
    return (t.data_ptr(), t._version, tuple(t.shape), t.dtype)


def _bn_affine(bn):
    """eval-mode BatchNorm1d as per-channel (scale, shift) f32, cached on the module."""
    ver = (_tensor_version(bn.running_mean), _tensor_version(bn.running_var),
           _tensor_version(bn.weight) if bn.weight is not None else None,
           _tensor_version(bn.bias) if bn.bias is not None else None, bn.eps)
    cached = getattr(bn, "_b2me_affine", None)
    if cached is not None and cached[0] == ver:
        return cached[1], cached[2]
    with torch.no_grad():
        inv = torch.rsqrt(bn.running_var.float() + bn.eps)
        scale = inv * bn.weight.float() if bn.weight is not None else inv
        shift = -bn.running_mean.float() * scale
        if bn.bias is not None:
            shift = shift + bn.bias.float()
        scale, shift = scale.contiguous(), shift.contiguous()
    bn._b2me_affine = (ver, scale, shift)
    return scale, shift


def _weight_f32(module):
    """[K, Cin, Cout] contiguous f32 view of the module's weights (conv kernel or nn.Linear)."""
    if hasattr(module, "kernel"):
        w = module.kernel
        ver = _tensor_version(w)
        cached = getattr(module, "_b2me_w", None)
        if cached is not None and cached[0] == ver:
            return cached[1]
        with torch.no_grad():
            w3 = w.detach().float()
            if w3.dim() == 2:
                w3 = w3.unsqueeze(0)
            w3 = w3.contiguous()
        module._b2me_w = (ver, w3)
        return w3
    lin = module.linear
    ver = _tensor_version(lin.weight)
    cached = getattr(module, "_b2me_w", None)
    if cached is not None and cached[0] == ver:
        return cached[1]
    with torch.no_grad():
        w3 = lin.weight.detach().float().t().contiguous().unsqueeze(0)  # [1, Cin, Cout]
    module._b2me_w = (ver, w3)
    return w3


def _weight_packed(module, K, Cin1, Cin2, Cout, op_code):
    """SWIZZLE_128B smem image of the weights for the tcgen05 kernel (bf16 or tf32 operands), cached per module."""
    w3 = _weight_f32(module)
    ver = (w3.data_ptr(), K, Cin1, Cin2, Cout, op_code)
    cache = getattr(module, "_b2me_packed", None)
    if cache is not None and cache[0] == ver and cache[2] is w3:
        return cache[1]
    nbytes = lib.b2me_tc_packed_bytes(K, Cin1, Cin2, Cout, op_code)
    packed = torch.empty((nbytes,), dtype=torch.uint8, device=w3.device)
    check(lib.b2me_tc_pack_weights(ptr(w3), K, Cin1, Cin2, Cout, op_code, ptr(packed), stream()), "tc_pack_weights")
    _count(1)
    module._b2me_packed = (ver, packed, w3)
    return packed


def _head_weights(lin):
    """second linear of a fused head: ([Chid, C2p] f32 zero-padded transpose of nn.Linear.weight, bias), cached."""
    ver = (_tensor_version(lin.weight), _tensor_version(lin.bias) if lin.bias is not None else None)
    cached = getattr(lin, "_b2me_head", None)
    if cached is not None and cached[0] == ver:
        return cached[1], cached[2]
    with torch.no_grad():
        C2, Chid = lin.weight.shape
        C2p = (C2 + 3) // 4 * 4
        w = torch.zeros((Chid, C2p), dtype=torch.float32, device=lin.weight.device)
        w[:, :C2] = lin.weight.detach().float().t()
        b = lin.bias.detach().float().contiguous() if lin.bias is not None else None
    lin._b2me_head = (ver, w.contiguous(), b)
    return lin._b2me_head[1], b


# ------------------------------------------------------------------------------------------------ execution
def run_pending(p: _Pending):
    if p.kind == "conv":
        return _run_conv(p)
    if p.kind == "elt":
        return _run_elt(p)
    raise AssertionError(p.kind)


def _run_elt(p):
    x = p.src[0]._materialize()
    V, C = x.shape
    res = p.residual._materialize() if p.residual is not None else None
    out_dtype = x.dtype
    out = torch.empty((V, C), dtype=out_dtype, device=x.device)
    check(lib.b2me_affine_act(ptr(x), dtype_code(x.dtype), V, C, ptr(p.scale), ptr(p.shift), ptr(res),
                              dtype_code(res.dtype) if res is not None else 0, p.act, p.slope, ptr(out),
                              dtype_code(out_dtype), stream()), "affine_act")
    _count(1)
    return out


def _profile_conv(kind, p, Cin, masks=None, extra=None):
    """bench.py hook: 'census' records the algorithmic work of every convolution launch (pairs = sum_k P_k) and, for
    the tcgen05 kernel, the (256-row tile pair, offset) passes it executes (zero-filled rows included),
    'events' brackets the launch with CUDA events on the launching stream."""
    prof = get_profile()
    if prof is None:
        return None
    if prof["mode"] == "census":
        pairs = int((p.nbr >= 0).sum().item()) if p.nbr is not None else int(p.V_out)
        rec = dict(kind=kind, K=p.K, Cin=Cin, Cout=p.Cout, V_out=int(p.V_out), pairs=pairs)
        if kind == "tc":
            tiles = (int(p.V_out) + 255) // 256
            if masks is not None:
                m = masks[:tiles].to(torch.int64) & 0xFFFFFFFF
                rec["passes"] = int(sum(int(((m >> k) & 1).sum().item()) for k in range(p.K)))
            else:
                rec["passes"] = tiles * p.K
        if extra:
            rec.update(extra)
        prof["records"].append(rec)
        return None
    ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
    ev[0].record()
    prof["records"].append(ev)
    return ev


def _tc_mode():
    """(torch dtype of the activations, operand dtype code, output dtype code) of the tcgen05 path, or None."""
    mode = _State.compute_mode
    if mode == "bf16":
        return torch.bfloat16, _lib.BF16, _lib.BF16
    if mode == "tf32":
        return torch.float32, _lib.TF32, _lib.TF32
    return None


def _run_conv(p):
    srcs = p.src
    f1 = srcs[0]._materialize()
    f2 = srcs[1]._materialize() if len(srcs) > 1 else None
    Cin1 = f1.shape[1]
    Cin2 = f2.shape[1] if f2 is not None else 0
    K, V_out, Cout = p.K, p.V_out, p.Cout
    dev = f1.device
    res = p.residual._materialize() if p.residual is not None else None
    small_out = p.extra.get("f32_out", False)
    tc = _tc_mode()
    use_tc = (tc is not None and not small_out and lib.b2me_tc_supported(K, Cin1, Cin2, Cout) == 1)
    if use_tc:
        adt, op_code, out_code = tc
        f1 = _as_dtype(f1, adt)
        if f2 is not None:
            f2 = _as_dtype(f2, adt)
        if res is not None:
            res = _as_dtype(res, adt)
        packed = _weight_packed(p.module, K, Cin1, Cin2, Cout, op_code)
        out = torch.empty((V_out, Cout), dtype=adt, device=dev)
        perm, masks = p.perm() if p.perm is not None else (None, None)
        ev = _profile_conv("tc", p, Cin1 + Cin2, masks)
        check(lib.b2me_spconv_fwd_tc(ptr(f1), Cin1, ptr(f2), Cin2, f1.shape[0], op_code, ptr(packed), ptr(p.nbr),
                                     ptr(perm), ptr(masks), K, V_out, Cout, ptr(p.scale), ptr(p.shift), ptr(res),
                                     p.act, p.slope, ptr(out), out_code, _State.tc_flags, stream()), "spconv_fwd_tc")
        if ev is not None:
            ev[1].record()
        _count(1)
        return out
    # SIMT fp32-accumulate path (exact fp32 mode; shapes the tensor-core kernel does not take, e.g. the Cin = 3 stem)
    if f2 is not None and f2.dtype != f1.dtype:
        f2 = _as_dtype(f2, f1.dtype)
    W = _weight_f32(p.module)
    mode = _State.compute_mode
    if small_out or mode == "f32":
        out_dtype, out_code = torch.float32, _lib.F32
    elif mode == "tf32":
        out_dtype, out_code = torch.float32, _lib.TF32  # feeds tf32 tensor-core layers: stored rounded to tf32
    else:
        out_dtype, out_code = torch.bfloat16, _lib.BF16
    out = torch.empty((V_out, Cout), dtype=out_dtype, device=dev)
    ev = _profile_conv("simt", p, Cin1 + Cin2)
    check(lib.b2me_spconv_fwd_simt(ptr(f1), Cin1, ptr(f2), Cin2, dtype_code(f1.dtype), ptr(W), ptr(p.nbr), K, V_out,
                                   Cout, ptr(p.scale), ptr(p.shift), ptr(res),
                                   dtype_code(res.dtype) if res is not None else 0, p.act, p.slope, ptr(out),
                                   out_code, stream()), "spconv_fwd_simt")
    if ev is not None:
        ev[1].record()
    _count(1)
    return out


def _try_fused_head(module, x):
    """MinkowskiLinear(Cin, Chid) -> activation -> this MinkowskiLinear(Chid, C2 <= 16) as one launch
    (b2me_head_fused_tc): x must still be the pending first linear. Returns the logits SparseTensor or None."""
    p = x._pending
    tc = _tc_mode()
    if (not _State.fuse_head or tc is None or x._F is not None or p is None or p.kind != "conv" or p.K != 1
            or p.nbr is not None or len(p.src) != 1 or p.residual is not None or p.extra.get("f32_out")
            or not hasattr(p.module, "linear")):
        return None
    lin2 = module.linear
    Chid, C2 = p.Cout, lin2.out_features
    f = p.src[0]._materialize()
    Cin = f.shape[1]
    if lin2.in_features != Chid or C2 > 16 or lib.b2me_tc_supported(1, Cin, 0, Chid) != 1:
        return None
    n_tile = Chid if Chid <= 256 else (256 if Chid % 256 == 0 else (192 if Chid % 192 == 0 else 128))
    if n_tile % 64 or Chid % n_tile:
        return None
    adt, op_code, _ = tc
    f = _as_dtype(f, adt)
    V = f.shape[0]
    packed = _weight_packed(p.module, 1, Cin, 0, Chid, op_code)
    W2p, b2 = _head_weights(lin2)
    logits = torch.empty((V, C2), dtype=torch.float32, device=f.device)
    amax = torch.empty((max(V, 1),), dtype=torch.uint8, device=f.device)
    ev = _profile_conv("tc", p, Cin, None, dict(fused_head=C2))
    check(lib.b2me_head_fused_tc(ptr(f), Cin, V, op_code, ptr(packed), Chid, ptr(p.scale), ptr(p.shift), p.act, p.slope,
                                 ptr(W2p), ptr(b2), C2, ptr(logits), ptr(amax), _State.tc_flags, stream()),
          "head_fused_tc")
    if ev is not None:
        ev[1].record()
    _count(1)
    res = x._child(features=logits)
    res._row_argmax = amax[:V]
    return res


# ------------------------------------------------------------------------------------------------ op builders
def _sources(x: SparseTensor):
    if x._F is None and x._cat is not None:
        return list(x._cat)
    return [x]


def conv_forward(module, x: SparseTensor):
    """MinkowskiConvolution / MinkowskiConvolutionTranspose forward -> lazy SparseTensor."""
    mgr = x.coordinate_manager
    key = x.coordinate_map_key
    ks, st = module.kernel_size, module.stride
    if module.dilation != 1:
        raise NotImplementedError("dilated sparse convolution is outside the hot-path subset")
    if not module.is_transpose:
        perm = None
        if ks == 1 and st == 1:
            out_key, nbr, K = key, None, 1
        elif ks == 3 and st == 1:
            out_key, nbr, K = key, mgr.kernel_map_k3(key), 27
            perm = lambda: mgr.perm_k3(key)  # noqa: E731
        elif ks == 2 and st == 2:
            out_key, rec = mgr.stride_down(key)
            nbr, K = rec["nbr_down"], 8
            perm = lambda: mgr.perm_stride(rec, "down")  # noqa: E731
        else:
            raise NotImplementedError(f"MinkowskiConvolution(kernel_size={ks}, stride={st}) is outside the subset "
                                      "used by MinkUNet (k3 s1, k2 s2, k1 s1)")
    else:
        if ks == 2 and st == 2:
            out_key, rec = mgr.stride_up(key)
            nbr, K = rec["nbr_up"], 8
            perm = lambda: mgr.perm_stride(rec, "up")  # noqa: E731
        else:
            raise NotImplementedError(f"MinkowskiConvolutionTranspose(kernel_size={ks}, stride={st})")
    V_out = mgr.level(out_key).V
    p = _Pending("conv", _sources(x), module=module, nbr=nbr, perm=perm, K=K, V_out=V_out, Cout=module.out_channels)
    if module.bias is not None:
        p.shift = module.bias.detach().float().reshape(-1).contiguous()
        p.stage = 1
    return SparseTensor(coordinate_map_key=out_key, coordinate_manager=mgr, _pending=p)


def linear_forward(module, x: SparseTensor):
    lin = module.linear
    Cout = lin.out_features
    if Cout <= 16:
        fused = _try_fused_head(module, x)
        if fused is not None:
            return fused
        # K5: small-N head, fp32 logits straight from the (bf16 or f32) voxel features
        F = x._materialize()
        V = F.shape[0]
        out = torch.empty((V, Cout), dtype=torch.float32, device=F.device)
        amax = torch.empty((max(V, 1),), dtype=torch.uint8, device=F.device)
        Wt = lin.weight.detach().float().contiguous()
        b = lin.bias.detach().float().contiguous() if lin.bias is not None else None
        check(lib.b2me_linear_small(ptr(F), dtype_code(F.dtype), V, F.shape[1], ptr(Wt), ptr(b), Cout, ptr(out),
                                    ptr(amax), stream()), "linear_small")
        _count(1)
        res = x._child(features=out)
        res._row_argmax = amax[:V]  # per-voxel arg-max of the logits (lowest index wins ties), free by-product
        return res
    p = _Pending("conv", _sources(x), module=module, nbr=None, K=1, V_out=x.num_rows, Cout=Cout)
    if lin.bias is not None:
        p.shift = lin.bias.detach().float().contiguous()
        p.stage = 1
    return x._child(_pending=p)


def affine_forward(x: SparseTensor, scale, shift):
    """eval BatchNorm as per-channel affine: folded into the producer when it has not been sealed yet."""
    p = x._pending
    if x._F is None and p is not None and p.stage <= 1:
        q = p.clone()
        if q.scale is None and q.shift is None:
            q.scale, q.shift = scale, shift
        else:
            s0 = q.scale if q.scale is not None else torch.ones_like(scale)
            b0 = q.shift if q.shift is not None else torch.zeros_like(shift)
            q.scale = (s0 * scale).contiguous()
            q.shift = (b0 * scale + shift).contiguous()
        q.stage = 1
        return x._child(_pending=q)
    q = _Pending("elt", [x], V_out=x.num_rows, Cout=x.num_channels, scale=scale, shift=shift, stage=1)
    return x._child(_pending=q)


def act_forward(x: SparseTensor, act, slope=0.0):
    p = x._pending
    if x._F is None and p is not None and p.stage <= 2:
        q = p.clone()
        q.act, q.slope, q.stage = act, slope, 3
        return x._child(_pending=q)
    q = _Pending("elt", [x], V_out=x.num_rows, Cout=x.num_channels, act=act, slope=slope, stage=3)
    return x._child(_pending=q)


def global_pool(x: SparseTensor, mode):
    F = x._materialize()
    mgr = x.coordinate_manager
    key = x.coordinate_map_key
    B = mgr.batch_size(key)
    V, C = F.shape
    out = torch.empty((B, C), dtype=torch.float32, device=F.device)
    check(lib.b2me_global_pool(ptr(F), dtype_code(F.dtype), ptr(mgr.coordinates(key)), V, C, B, mode, ptr(out),
                               stream()), "global_pool")
    _count(1)
    # pooled tensor: one row per batch index at the origin
    from .core import _Level
    gkey = CoordinateMapKey(0, key._tag + "/global")
    if gkey not in mgr.levels:
        gc = torch.zeros((B, 4), dtype=torch.int32, device=F.device)
        gc[:, 0] = torch.arange(B, dtype=torch.int32, device=F.device)
        lv = _Level(gc, None, B)
        lv.batch_size = B
        mgr.levels[gkey] = lv
    return SparseTensor(features=out, coordinate_map_key=gkey, coordinate_manager=mgr)
