"""ctypes binding of libb2me.so (include/b2me.h). No torch types cross this boundary: raw device
pointers, sizes and the caller's cudaStream_t only. There is NO CPU fallback: if the library is missing
the import fails loudly with the build command to run."""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# B2ME_LIB_PATH: developer override used by tools/conv_probe.py to load an instrumented build (lib_debug/)
LIB_PATH = os.environ.get("B2ME_LIB_PATH") or os.path.join(os.path.dirname(_HERE), "lib", "libb2me.so")

F32, BF16, TF32 = 0, 1, 2
TC_FLAG_TMA, TC_FLAG_NO_ROT128, TC_FLAG_PF_BULK, TC_FLAG_PF_NONE, TC_FLAG_PF_NEAR = 1, 2, 4, 8, 16
ACT_NONE, ACT_RELU, ACT_LEAKY = 0, 1, 2

_vp, _i32, _i64, _sz, _f32, _f64 = C.c_void_p, C.c_int, C.c_int64, C.c_size_t, C.c_float, C.c_double

# name -> (restype, argtypes); mirrors include/b2me.h one to one
SIGNATURES = {
    "b2me_version": (_i32, []),
    "b2me_strerror": (C.c_char_p, [_i32]),
    "b2me_table_slots": (_i64, [_i64]),
    "b2me_table_bytes": (_sz, [_i64]),
    "b2me_unique_workspace_bytes": (_sz, [_i64, _i32]),
    "b2me_quantize_unique": (_i32, [_vp, _vp, _i64, _vp, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp, _sz, _vp]),
    "b2me_quantize_labels": (_i32, [_vp, _vp, _vp, _i64, _i64, _i32, _vp, _vp]),
    "b2me_stride_map": (_i32, [_vp, _i64, _i32, _vp, _vp, _vp, _vp, _vp, _sz, _vp, _sz, _vp]),
    "b2me_stride_kernel_maps": (_i32, [_vp, _vp, _i64, _i64, _vp, _vp, _vp]),
    "b2me_kernel_map_k3": (_i32, [_vp, _i64, _i32, _vp, _sz, _vp, _vp, _vp]),
    "b2me_block_rows": (_i32, [_vp, _i64, _i32, _vp, _vp, _i64, _vp, _vp]),
    "b2me_kernel_map_k3_blocks": (_i32, [_vp, _i64, _i32, _vp, _sz, _vp, _vp, _vp, _vp, _vp]),
    "b2me_row_masks": (_i32, [_vp, _i64, _i32, _vp, _vp, _vp]),
    "b2me_mask_sort_keys_rows": (_i32, [_vp, _vp, _i64, _i32, _vp, _vp]),
    "b2me_tile_masks_rows": (_i32, [_vp, _vp, _i64, _vp, _vp]),
    "b2me_mask_sort_keys": (_i32, [_vp, _i64, _i32, _vp, _vp, _sz, _vp]),
    "b2me_mask_sort_keys2_ws_bytes": (_sz, [_i64]),
    "b2me_mask_sort_keys2": (_i32, [_vp, _i64, _i32, _vp, _vp, _sz, _vp]),
    "b2me_mask_sort_keys_morton": (_i32, [_vp, _vp, _i64, _i32, _i32, _vp, _vp, _sz, _vp]),
    "b2me_mask_sort_keys64": (_i32, [_vp, _i64, _i32, _i32, _vp, _vp, _sz, _vp]),
    "b2me_spconv_fwd_simt": (_i32, [_vp, _i32, _vp, _i32, _i32, _vp, _vp, _i32, _i64, _i32, _vp, _vp, _vp, _i32,
                                    _i32, _f32, _vp, _i32, _vp]),
    "b2me_tc_packed_bytes": (_sz, [_i32, _i32, _i32, _i32, _i32]),
    "b2me_tc_supported": (_i32, [_i32, _i32, _i32, _i32]),
    "b2me_tc_pack_weights": (_i32, [_vp, _i32, _i32, _i32, _i32, _i32, _vp, _vp]),
    "b2me_tc_tile_masks": (_i32, [_vp, _vp, _i64, _i32, _vp, _vp]),
    "b2me_spconv_fwd_tc": (_i32, [_vp, _i32, _vp, _i32, _i64, _i32, _vp, _vp, _vp, _vp, _i32, _i64, _i32, _vp, _vp, _vp,
                                  _i32, _f32, _vp, _i32, _i32, _vp]),
    "b2me_head_fused_tc": (_i32, [_vp, _i32, _i64, _i32, _vp, _i32, _vp, _vp, _i32, _f32, _vp, _vp, _i32, _vp, _vp, _i32,
                                  _vp]),
    "b2me_affine_act": (_i32, [_vp, _i32, _i64, _i32, _vp, _vp, _vp, _i32, _i32, _f32, _vp, _i32, _vp]),
    "b2me_convert": (_i32, [_vp, _i32, _vp, _i32, _i64, _vp]),
    "b2me_linear_small": (_i32, [_vp, _i32, _i64, _i32, _vp, _vp, _i32, _vp, _vp, _vp]),
    "b2me_gather_rows": (_i32, [_vp, _i32, _i32, _vp, _i64, _vp, _vp]),
    "b2me_gather_labels": (_i32, [_vp, _vp, _i64, _vp, _vp]),
    "b2me_global_pool": (_i32, [_vp, _i32, _vp, _i64, _i32, _i32, _i32, _vp, _vp]),
    "b2me_keypoint_reduce": (_i32, [_vp, _i32, _vp, _i32, _vp, _vp, _vp]),
    "b2me_vote_center": (_i32, [_vp, _i32, _i32, _vp, _vp, _i32, _i32, _vp, _vp]),
    "b2me_translation_magic": (_i32, [_vp, _vp, _i32, _vp, _f32, _vp, _vp]),
    "b2me_sanity_check": (_i32, [_vp, _vp, _i32, _vp, _vp, _vp, _i32, _f32, _i32, _f64, _vp, _vp]),
    "b2me_cluster_workspace_bytes": (_sz, [_i64, _i32]),
    "b2me_largest_cluster": (_i32, [_vp, _vp, _i32, _i64, _f64, _vp, _vp, _vp, _sz, _vp]),
    "b2me_kabsch_batched": (_i32, [_vp, _vp, _vp, _i32, _i32, _vp, _vp, _vp]),
    "b2me_ingest_workspace_bytes": (_sz, [_i64]),
    "b2me_ingest_clouds": (_i32, [_vp, _i64, _vp, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "b2me_select_workspace_bytes": (_sz, [_i64]),
    "b2me_select_rows": (_i32, [_vp, _i32, _vp, _i64, _vp, _i32, _vp, _vp, _vp, _sz, _vp]),
    "b2me_gather_crops": (_i32, [_vp, _vp, _vp, _i64, _vp, _i32, _vp, _vp, _vp, _vp]),
    "b2me_center_segments": (_i32, [_vp, _vp, _i32, _vp, _vp, _vp]),
    "b2me_normalize_colors": (_i32, [_vp, _vp, _i64, _vp, _i32, _vp, _vp, _sz, _vp]),
    "b2me_fps": (_i32, [_vp, _i32, _i32, _i32, _vp, _vp, _vp]),
    "b2me_ball_query": (_i32, [_vp, _vp, _i32, _i32, _i32, _f32, _i32, _vp, _vp]),
    "b2me_three_nn": (_i32, [_vp, _vp, _i32, _i32, _i32, _vp, _vp, _vp]),
    "b2me_icp_workspace_bytes": (_sz, [_i64, _i32, _i32]),
    "b2me_icp_p2p_batched": (_i32, [_vp, _i32, _vp, _vp, _i32, _i64, _vp, _f64, _i32, _f64, _f64, _vp, _vp, _vp, _sz,
                                    _i32, _vp]),
}


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"libb2me.so not found at {LIB_PATH}. Build it with "
            f"`python {os.path.join(os.path.dirname(_HERE), 'build.py')}` (nvcc, sm_100a). "
            "This package has no CPU or PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing: fail loudly
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()


class B2MEError(RuntimeError):
    pass


def check(rc, what=""):
    if rc != 0:
        raise B2MEError(f"libb2me {what}: {lib.b2me_strerror(rc).decode()} (code {rc})")


def ptr(t):
    """device pointer of a tensor (None -> NULL)."""
    if t is None:
        return None
    return C.c_void_p(t.data_ptr())


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def dtype_code(dt):
    if dt == torch.float32:
        return F32
    if dt == torch.bfloat16:
        return BF16
    raise B2MEError(f"unsupported feature dtype {dt}")


def require_cuda(t, name):
    if not t.is_cuda:
        raise B2MEError(f"{name} must live on a CUDA device: this MinkowskiEngine build runs on B200 only "
                        "(no CPU fallback)")
    return t
