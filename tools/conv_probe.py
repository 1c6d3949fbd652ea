#!/usr/bin/env python
"""Stand-alone probe of the tcgen05 sparse convolution (K4b) on a realistic kernel map: times one layer shape on
the 5 mm voxel map of a few synthetic frames and, when the library was built with -DB2ME_TC_PROFILE
(B2ME_EXTRA_NVCC_FLAGS=-DB2ME_TC_PROFILE python build.py; B2ME_LIB_PATH=.../lib_debug/libb2me.so), prints where
each warp role of cluster 0 spent its cycles. Developer tool, not part of the product path."""
import argparse
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "markerless-robot-camera-calibration_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=4)
    ap.add_argument("--scale", type=float, default=200.0)
    ap.add_argument("--shapes", default="27:384:384,1:416:384,27:32:32,8:384:384,1:256:1024")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--level", type=int, default=1, help="tensor stride of the map (1 or 2)")
    ap.add_argument("--out", default=None)
    ap.add_argument("--path", default="tma", choices=["cpasync", "tma"], help="operand path (B2ME_TC_FLAG_TMA)")
    ap.add_argument("--no-rot128", action="store_true", help="384-column tiles: single accumulator (round-1 layout)")
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "tf32"])
    ap.add_argument("--prefetch", default="none", choices=["none", "near", "bulk"], help="L2 prefetch scheme")
    ap.add_argument("--sb", type=int, default=0, help="B-ring stages override (flags bits 8-10; 0 = library default)")
    ap.add_argument("--rounds", type=int, default=5, help="--sweep: interleaved timing rounds per variant")
    ap.add_argument("--lib-b", default=None,
                    help="--sweep: a second build of libb2me.so; every variant is also timed through it (suffix @B): "
                         "same-box, interleaved A/B of two kernel versions")
    ap.add_argument("--sweep", action="store_true",
                    help="time every shape under a list of flag variants (operand path, prefetch mode, B-ring depth, "
                         "accumulator layout) on the same map: same-box A/B")
    ap.add_argument("--head", type=int, default=0, help="K=1 shapes: time the fused head with this many classes instead")
    a = ap.parse_args()
    import MinkowskiEngine as ME
    from MinkowskiEngine._lib import (lib, ptr, stream, check, BF16, TF32, TC_FLAG_TMA, TC_FLAG_NO_ROT128,
                                      TC_FLAG_PF_BULK, TC_FLAG_PF_NONE, TC_FLAG_PF_NEAR)
    op = TF32 if a.dtype == "tf32" else BF16
    flags = ((TC_FLAG_TMA if a.path == "tma" else 0) | (TC_FLAG_NO_ROT128 if a.no_rot128 else 0)
             | {"near": TC_FLAG_PF_NEAR, "bulk": TC_FLAG_PF_BULK, "none": TC_FLAG_PF_NONE}[a.prefetch]
             | ((a.sb & 7) << 8))
    variants = [("default", flags)]
    if a.sweep:
        # "default" = the package default (TMA operand path, rot128 accumulators, no L2 prefetch)
        T = TC_FLAG_TMA
        variants = [("default", T), ("cpasync", 0), ("tma_no_rot128", T | TC_FLAG_NO_ROT128), ("tma_sb2", T | (2 << 8)),
                    ("tma_sb4", T | (4 << 8)), ("cpasync_pf_near", TC_FLAG_PF_NEAR)]
    adt = torch.float32 if a.dtype == "tf32" else torch.bfloat16
    from b200calib.synthetic import make_frame
    frames = [make_frame(13000 + i) for i in range(a.frames)]
    pts = torch.from_numpy(np.concatenate([f["points"] for f in frames])).cuda()
    bidx = torch.from_numpy(np.concatenate([np.full(len(f["points"]), i, np.float32) for i, f in enumerate(frames)])).cuda()
    coords = torch.cat((bidx.unsqueeze(1), pts * a.scale), 1)
    fld = ME.TensorField(features=torch.rand(len(pts), 3, device="cuda"), coordinates=coords, device="cuda")
    sp = fld.sparse()
    mgr, key = sp.coordinate_manager, sp.coordinate_map_key
    if a.level == 2:
        key, _ = mgr.stride_down(key)
    V = mgr.level(key).V
    nbr27 = mgr.kernel_map_k3(key)
    perm27, masks27 = mgr.perm_k3(key)
    has_prof = hasattr(lib, "b2me_tc_prof_read") if False else True
    try:
        prof_read = C.CDLL(lib._name).b2me_tc_prof_read
        prof_read.restype = C.c_int
        prof_read.argtypes = [C.c_void_p, C.c_int]
    except AttributeError:
        prof_read = None
    res = []
    lib_a, lib_b = lib, None
    if a.lib_b:
        from MinkowskiEngine._lib import SIGNATURES
        lib_b = C.CDLL(a.lib_b)
        for name, (rt, at) in SIGNATURES.items():
            fn = getattr(lib_b, name)
            fn.restype, fn.argtypes = rt, at
        variants = [(n, f, lib_a) for n, f in variants] + [(n + "@B", f, lib_b) for n, f in variants]
    else:
        variants = [(n, f, lib_a) for n, f in variants]
    for spec in a.shapes.split(","):
        K, cin, cout = [int(x) for x in spec.split(":")]
        g = torch.Generator(device="cuda").manual_seed(1)
        x = torch.randn(V, cin, device="cuda", generator=g).to(adt)
        W = (torch.randn(K, cin, cout, device="cuda", generator=g) / np.sqrt(K * cin)).contiguous()
        packed = torch.empty(lib.b2me_tc_packed_bytes(K, cin, 0, cout, op), dtype=torch.uint8, device="cuda")
        check(lib.b2me_tc_pack_weights(ptr(W), K, cin, 0, cout, op, ptr(packed), stream()))
        if K == 27:
            nbr, perm, masks = nbr27, perm27, masks27
        elif K == 8:
            ck, rec = mgr.stride_down(key)
            nbr = rec["nbr_up"]
            perm, masks = mgr.perm_stride(rec, "up")
        else:
            nbr, perm, masks = None, None, None
        out = torch.empty(V, cout, device="cuda", dtype=adt)
        head = a.head if (K == 1 and a.head) else 0
        if head:
            W2p = torch.randn(cout, (head + 3) // 4 * 4, device="cuda", generator=g)
            logits = torch.empty(V, head, device="cuda")
            amax = torch.empty(V, dtype=torch.uint8, device="cuda")
        pairs = int((nbr >= 0).sum()) if nbr is not None else V

        def run():
            if head:
                check(lib.b2me_head_fused_tc(ptr(x), cin, V, op, ptr(packed), cout, None, None, 2, 0.01, ptr(W2p), None,
                                             head, ptr(logits), ptr(amax), flags, stream()))
                return
            check(lib.b2me_spconv_fwd_tc(ptr(x), cin, None, 0, x.shape[0], op, ptr(packed), ptr(nbr), ptr(perm),
                                         ptr(masks), K, V, cout, None, None, None, 1, 0.0, ptr(out), op, flags,
                                         stream()))
        if a.sweep:
            # interleaved rounds (the SM clock drifts under the power cap): every variant is timed once per round, the
            # reported time is the median over the rounds
            ref_out = None
            times = {vn: [] for vn, _, _ in variants}
            same = {}
            for rnd in range(a.rounds):
                for vname, vflags, vlib in variants:
                    flags, lib = vflags, vlib
                    if rnd == 0:
                        run()
                        torch.cuda.synchronize()
                        cur = (logits if head else out).clone()
                        same[vname] = True if ref_out is None else bool(torch.equal(cur, ref_out))
                        ref_out = cur if ref_out is None else ref_out
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    for _ in range(a.reps):
                        run()
                    e1.record()
                    torch.cuda.synchronize()
                    times[vname].append(e0.elapsed_time(e1) / a.reps)
            lib = lib_a
            for vname, _, _ in variants:
                ms = float(np.median(times[vname]))
                print(f"K={K} {cin}->{cout} V={V} [{vname:14s}] median {ms:8.3f} ms (min {min(times[vname]):.3f})  "
                      f"{2.0 * pairs * cin * cout / ms / 1e9:7.0f} TFLOP/s  bit-identical to the first variant: {same[vname]}",
                      flush=True)
                res.append(dict(K=K, Cin=cin, Cout=cout, V=V, variant=vname, ms=ms, ms_min=min(times[vname]),
                                same=same[vname]))
            continue
        run()
        torch.cuda.synchronize()
        if prof_read:
            prof_read(None, 1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.reps):
            run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.reps
        rec = dict(K=K, Cin=cin, Cout=cout, V=V, pairs=pairs, ms=ms, tflops=2.0 * pairs * cin * cout / ms / 1e9)
        if prof_read:
            buf = (C.c_ulonglong * 128)()
            prof_read(buf, 1)
            prof = np.array(list(buf), dtype=np.float64).reshape(2, 8, 8) / a.reps
            rec["prof"] = prof.tolist()
            names = ["producer", "mma", "bload", "prefetch", "epilogue", "relay"]
            print(f"--- K={K} {cin}->{cout} V={V}: {ms:.3f} ms, {rec['tflops']:.0f} TFLOP/s")
            for r in range(2):
                for role in (0, 1, 4, 5):  # producer / mma / epilogue: per item or tile; counters c1.. = waits
                    pr = prof[r, role]
                    if pr.sum() == 0:
                        continue
                    n = max(pr[7], 1)
                    print(f"  rank{r} {names[role]:9s} total {pr[0]:12.0f} cyc  n={pr[7]:9.0f}  "
                          + " ".join(f"c{q}={pr[q] / n:8.1f}" for q in range(1, 7)) + "  (cycles per item/tile)")
        else:
            print(f"K={K} {cin}->{cout} V={V}: {ms:.3f} ms, {rec['tflops']:.0f} TFLOP/s")
        res.append(rec)
    if a.out:
        json.dump(res, open(a.out, "w"))


if __name__ == "__main__":
    main()
