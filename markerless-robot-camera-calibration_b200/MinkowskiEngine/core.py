"""Coordinate manager, SparseTensor, TensorField and the lazy fused-op record of the B200
MinkowskiEngine-compatible package.

Design: reference model code calls conv -> batch-norm -> (+= residual) -> relu as separate
modules (model/backbone/minkunet.py:126-181). Each of those returns a SparseTensor whose features are a
*pending* fused operation; the chain is executed as ONE kernel launch (gather-GEMM with the affine,
residual and activation in its epilogue) when the features are first needed. ME.cat of two tensors is a
lazy two-source view that the convolution kernel consumes without materialising the concatenation.
"""
import threading
from dataclasses import dataclass, field as dc_field
from enum import Enum
from typing import List, Optional

import torch

from . import _lib
from ._lib import lib, check, ptr, stream, dtype_code


# ------------------------------------------------------------------------------------------------ enums
class SparseTensorQuantizationMode(Enum):
    RANDOM_SUBSAMPLE = 0
    UNWEIGHTED_AVERAGE = 1
    UNWEIGHTED_SUM = 2
    NO_QUANTIZATION = 3
    MAX_POOL = 4
    SPLAT_LINEAR_INTERPOLATION = 5


class MinkowskiAlgorithm(Enum):
    DEFAULT = 0
    MEMORY_EFFICIENT = 1
    SPEED_OPTIMIZED = 2


class SparseTensorOperationMode(Enum):
    SEPARATE_COORDINATE_MANAGER = 0
    SHARE_COORDINATE_MANAGER = 1


class _State:
    mask_sort = True  # tcgen05 convolutions take a neighbour-mask-sorted row permutation (tile-level offset skipping)
    mask_sort_min_rows = 32768  # smaller maps keep their natural row order (the sort costs more than it saves)
    mask_sort_block = 0  # rows per locality block of that sort, 0 = whole map (measured: 0 is fastest)
    # K = 27 maps: per-segment order of the common offsets (b2me_mask_sort_keys2). Off by default: at 32 frames the
    # one-level order already reaches 0.885 row efficiency (two-level 0.901: conv -1 ms, key kernels +5 ms per step)
    mask_sort_two_level = False
    # k3 maps: Morton order inside a mask group (b2me_mask_sort_keys_morton). Off by default: measured on one box, the
    # convolution rate does not change (MMA-executed 1053-1060 vs 1058-1068 TFLOP/s) and the 64-bit sort costs ~2 ms
    mask_sort_morton = False
    k3_block_min_rows = 65536  # k3 kernel maps of at least this many rows go through 4x4x4 blocks (b2me_kernel_map_k3_blocks)
    compute_dtype = torch.float32
    # "bf16" activations on tcgen05 | "tf32": fp32 activations holding tf32 values on tcgen05 kind::tf32 | "f32": SIMT
    compute_mode = "f32"
    # B2ME_TC_FLAG_* passed to every tcgen05 launch (operand path, accumulator layout). Default operand path: TMA
    # (tile::gather4 rows + 2-D weight boxes) - interleaved A/B medians of round 2 (profiles/r02_ab_medians.md): 3-9 %
    # faster than the cp.async path on the K = 27 384-channel layers, 10-14 % on K = 1 416->384, 22-26 % on the narrow
    # (32 / 128 channel) layers, never slower; bit-identical results, both paths under the same parity tests
    tc_flags = _lib.TC_FLAG_TMA
    fuse_head = True  # MinkowskiLinear(.., hidden) -> act -> MinkowskiLinear(hidden, C <= 16) as ONE launch
    launches = 0  # kernels launched through libb2me since the last reset (bench.py reads this)
    lock = threading.Lock()  # several host threads may drive the library (one stream each, pipeline.predict_stream)


_tls = threading.local()  # per-thread profile hook (bench.py), see ops._profile_conv


def set_compute_dtype(dtype):
    """torch.float32 / "f32": exact-fp32 SIMT convolutions.
    torch.bfloat16 / "bf16": bf16 activations, tcgen05 convolutions with fp32 accumulation (BASELINE.json config 2).
    "tf32": fp32 activations, tcgen05 kind::tf32 convolutions with fp32 accumulation: the tensor-core mode that
    meets the 1e-3 fp32 tolerance (activations and weights are rounded to tf32, 10-bit mantissa, to nearest)."""
    if dtype in (torch.float32, "f32", "fp32", "float32"):
        _State.compute_dtype, _State.compute_mode = torch.float32, "f32"
    elif dtype in (torch.bfloat16, "bf16", "bfloat16"):
        _State.compute_dtype, _State.compute_mode = torch.bfloat16, "bf16"
    elif dtype in ("tf32", "tfloat32"):
        _State.compute_dtype, _State.compute_mode = torch.float32, "tf32"
    else:
        raise ValueError("compute dtype must be torch.float32, torch.bfloat16 or 'tf32'")


def get_compute_dtype():
    return _State.compute_dtype


def get_compute_mode():
    """'f32' | 'bf16' | 'tf32'."""
    return _State.compute_mode


def set_tc_operand_path(path):
    """how the tcgen05 convolution fetches its operands: "tma" (default: cp.async.bulk.tensor tile::gather4 rows + 2-D
    weight boxes) or "cpasync" (16-byte cp.async gathers + cp.async.bulk weights). Bit-identical results."""
    if path not in ("cpasync", "tma"):
        raise ValueError('operand path must be "cpasync" or "tma"')
    _State.tc_flags = (_State.tc_flags & ~_lib.TC_FLAG_TMA) | (_lib.TC_FLAG_TMA if path == "tma" else 0)


def get_tc_operand_path():
    return "tma" if _State.tc_flags & _lib.TC_FLAG_TMA else "cpasync"


def set_tc_rot128(on):
    """384-column tiles: early release of the 256-column accumulator part + alternating 128-column regions (True,
    default) or the single 384-column accumulator of round 1 (False). Bit-identical results."""
    _State.tc_flags = (_State.tc_flags & ~_lib.TC_FLAG_NO_ROT128) | (0 if on else _lib.TC_FLAG_NO_ROT128)


def set_tc_prefetch(mode):
    """L2 prefetch of the next kernel offset's gathered rows in the tcgen05 convolution: "none" (default), "near"
    (one prefetch.global.L2 per 128-byte chunk) or "bulk" (one cp.async.bulk.prefetch.L2 per row)."""
    bits = {"none": 0, "near": _lib.TC_FLAG_PF_NEAR, "bulk": _lib.TC_FLAG_PF_BULK}[mode]
    _State.tc_flags = (_State.tc_flags & ~(_lib.TC_FLAG_PF_BULK | _lib.TC_FLAG_PF_NONE | _lib.TC_FLAG_PF_NEAR)) | bits


def set_fuse_head(on):
    """fused 256 -> 1024 -> C head in one launch (True, default) or the two-launch path with the hidden activation in
    HBM (False; A/B and parity tests)."""
    _State.fuse_head = bool(on)


def reset_launch_count():
    _State.launches = 0


def launch_count():
    return _State.launches


def set_mask_sort(flag):
    """enable / disable the mask-sorted row permutation of the tcgen05 convolutions (results are bit-identical
    either way; only the number of (tile, offset) MMA passes changes)."""
    _State.mask_sort = bool(flag)


def set_mask_sort_morton(on):
    """A/B switch: Morton (True) or first-occurrence (False, default) order inside a mask group of the k3 maps."""
    _State.mask_sort_morton = bool(on)


def set_mask_sort_two_level(on):
    """A/B switch: two-level (True) or one-level (False, default) mask-sort keys for the K = 27 maps."""
    _State.mask_sort_two_level = bool(on)


def set_k3_block_min_rows(rows):
    """k3 kernel maps with at least `rows` rows are built through the 4 x 4 x 4 block arrays of the map two stride-2
    levels up (1..8 probes of a small table per voxel instead of 26 of the voxel table); smaller ones, and maps whose
    coarser levels cannot be formed, use the direct probe kernel. Identical maps either way."""
    _State.k3_block_min_rows = int(rows)


def set_mask_sort_block(rows):
    """rows of one locality block of the mask sort (0 = sort the whole map by mask only)."""
    _State.mask_sort_block = int(rows)


def mask_sorted_perm(nbr, V, K, block_rows=None, coords=None, ts=1, row_masks=None, offset_counts=None):
    """row order that groups rows with the same neighbour pattern inside blocks of `block_rows` consecutive rows
    (K3b keys + a device sort). With row_masks / offset_counts (by-products of the kernel-map pass) the keys come from
    4 bytes per row instead of the whole nbr table."""
    if block_rows is None:
        block_rows = _State.mask_sort_block
    ws = torch.empty((128,), dtype=torch.uint8, device=nbr.device)
    if block_rows == 0 and coords is not None and _State.mask_sort_morton:
        # mask keys + Morton tie-break: rows with the same neighbour pattern follow a space-filling curve
        keys = torch.empty((max(V, 1),), dtype=torch.int64, device=nbr.device)
        check(lib.b2me_mask_sort_keys_morton(ptr(nbr), ptr(coords), V, K, int(ts), ptr(keys), ptr(ws), ws.numel(),
                                             stream()), "mask_sort_keys_morton")
        _count(2)
    elif block_rows == 0 and _State.mask_sort_two_level and K == 27:
        # two-level keys: global rarest-first segment + per-segment order of the remaining offsets
        ws = torch.empty((lib.b2me_mask_sort_keys2_ws_bytes(V),), dtype=torch.uint8, device=nbr.device)
        keys = torch.empty((max(V, 1),), dtype=torch.int32, device=nbr.device)
        check(lib.b2me_mask_sort_keys2(ptr(nbr), V, K, ptr(keys), ptr(ws), ws.numel(), stream()), "mask_sort_keys2")
        _count(4)
    elif block_rows == 0 and row_masks is not None:
        keys = torch.empty((max(V, 1),), dtype=torch.int32, device=nbr.device)
        check(lib.b2me_mask_sort_keys_rows(ptr(row_masks), ptr(offset_counts), V, K, ptr(keys), stream()),
              "mask_sort_keys_rows")
        _count(1)
    elif block_rows == 0:  # mask-only keys fit 32 bits: half the radix passes of the device sort
        keys = torch.empty((max(V, 1),), dtype=torch.int32, device=nbr.device)
        check(lib.b2me_mask_sort_keys(ptr(nbr), V, K, ptr(keys), ptr(ws), ws.numel(), stream()), "mask_sort_keys")
        _count(2)
    else:
        keys = torch.empty((max(V, 1),), dtype=torch.int64, device=nbr.device)
        check(lib.b2me_mask_sort_keys64(ptr(nbr), V, K, block_rows, ptr(keys), ptr(ws), ws.numel(), stream()),
              "mask_sort_keys64")
        _count(2)
    return torch.sort(keys[:V])[1].to(torch.int32)


def tile_masks(nbr, perm, V, K, row_masks=None):
    """per 256-row tile pair: the kernel offsets any of its rows needs (consumed by b2me_spconv_fwd_tc)."""
    masks = torch.empty(((max(V, 1) + 255) // 256,), dtype=torch.int32, device=nbr.device)
    if row_masks is not None:
        check(lib.b2me_tile_masks_rows(ptr(row_masks), ptr(perm), V, ptr(masks), stream()), "tile_masks_rows")
    else:
        check(lib.b2me_tc_tile_masks(ptr(nbr), ptr(perm), V, K, ptr(masks), stream()), "tc_tile_masks")
    _count(1)
    return masks


def set_profile(mode):
    """None | 'census' | 'events' -> the record list that ops._profile_conv appends to (state of the CALLING thread:
    every host thread of a pipelined run collects the launches of its own batches)."""
    _tls.profile = None if mode is None else dict(mode=mode, records=[])
    return None if mode is None else _tls.profile["records"]


def get_profile():
    return getattr(_tls, "profile", None)


def _count(n=1):
    with _State.lock:
        _State.launches += n


# ------------------------------------------------------------------------------------------------ coordinate maps
class CoordinateMapKey:
    def __init__(self, tensor_stride, tag=""):
        self._ts = int(tensor_stride)
        self._tag = tag

    def get_tensor_stride(self):
        return [self._ts] * 3

    def get_key(self):
        return ([self._ts] * 3, self._tag)

    def __eq__(self, other):
        return isinstance(other, CoordinateMapKey) and self._ts == other._ts and self._tag == other._tag

    def __hash__(self):
        return hash((self._ts, self._tag))

    def __repr__(self):
        return f"CoordinateMapKey(tensor_stride={self._ts}, tag={self._tag!r})"


class _Level:
    """One coordinate map: rows in first-occurrence order + its hash table (key -> row)."""

    def __init__(self, coords, table, V):
        self.coords = coords  # [V,4] int32 device
        self.table = table    # uint8 device buffer (b2me_table_bytes)
        self.V = V
        self.nbr_k3 = None
        self.masks_k3 = None  # (row occupancy masks [V] i32, per-offset neighbour counts [32] i32) of the k3 map
        self.perm_k3 = None
        self.down = None      # dict(in2out, koff, nbr_down, nbr_up, coarse_key)
        self.batch_size = None


class CoordinateManager:
    """Caches coordinate maps and kernel maps per (tensor stride, tag), like ME's manager."""

    def __init__(self, D=3, device=None):
        self.D = D
        self.device = device
        self.levels = {}
        self.field_inverse = None  # TensorField point -> voxel row (int32)
        self.number_of_batches = None

    # -- creation of the stride-1 map from raw coordinates
    def insert_coordinates(self, coords, feats, mode, tag=""):
        """coords [N,4] float32 (floor is applied) or int32; feats [N,C] float32 or None.
        Returns (key, voxel_feats [V,C] f32 or None, inverse [N] i32, first_idx [V] i32)."""
        dev = coords.device
        N = coords.shape[0]
        Cf = 0 if feats is None else feats.shape[1]
        is_float = coords.dtype.is_floating_point
        coords = coords.contiguous().to(torch.float32 if is_float else torch.int32)
        if feats is not None:
            feats = feats.contiguous().to(torch.float32)
        cap = max(N, 1)
        out_coords = torch.empty((cap, 4), dtype=torch.int32, device=dev)
        out_feats = torch.empty((cap, max(Cf, 1)), dtype=torch.float32, device=dev) if Cf else None
        inverse = torch.empty((cap,), dtype=torch.int32, device=dev)
        first_idx = torch.empty((cap,), dtype=torch.int32, device=dev)
        counts = torch.zeros((4,), dtype=torch.int32, device=dev)
        table = torch.empty((lib.b2me_table_bytes(N),), dtype=torch.uint8, device=dev)
        ws = torch.empty((lib.b2me_unique_workspace_bytes(N, Cf),), dtype=torch.uint8, device=dev)
        check(lib.b2me_quantize_unique(ptr(coords) if is_float else None, None if is_float else ptr(coords), N,
                                       ptr(feats), Cf, mode, ptr(out_coords), ptr(out_feats), ptr(inverse),
                                       ptr(first_idx), ptr(counts), ptr(table), table.numel(), ptr(ws), ws.numel(),
                                       stream()), "quantize_unique")
        _count(8)
        V, err = counts[:2].tolist()  # the one host sync of voxelisation
        if err:
            raise _lib.B2MEError("voxel coordinate outside the +-131072 key range, batch index >= 1024, or a "
                                 f"feature value beyond the fixed-point range (flag {err})")
        key = CoordinateMapKey(1, tag)
        self.levels[key] = _Level(out_coords[:V], table, V)
        return key, (out_feats[:V] if Cf else None), inverse[:N], first_idx[:V]

    def level(self, key):
        return self.levels[key]

    def coordinates(self, key):
        return self.levels[key].coords

    # -- K3
    def _block_parents(self, key):
        """(in2out ts -> 2 ts, in2out 2 ts -> 4 ts, level at 4 ts) when both stride-2 maps above `key` can be formed."""
        try:
            k1, rec1 = self.stride_down(key)
            k2, rec2 = self.stride_down(k1)
        except _lib.B2MEError:
            return None
        lv2 = self.levels[k2]
        if lv2.table is None or lv2.V == 0:
            return None
        return rec1["in2out"], rec2["in2out"], lv2

    def kernel_map_k3(self, key):
        lv = self.levels[key]
        if lv.nbr_k3 is None:
            dev = lv.coords.device
            nbr = torch.empty((max(lv.V, 1), 27), dtype=torch.int32, device=dev)
            row_masks = torch.empty((max(lv.V, 1),), dtype=torch.int32, device=dev)
            counts = torch.empty((32,), dtype=torch.int32, device=dev)
            parents = self._block_parents(key) if lv.V >= _State.k3_block_min_rows else None
            if parents is not None:
                # large map: through the 4 x 4 x 4 blocks = the rows of the map two stride-2 levels up (the UNet's own
                # coordinate hierarchy; those maps are cached for the stride-2 convolutions that follow)
                in2out1, in2out2, lv2 = parents
                brows = torch.empty((max(lv2.V, 1) * 64,), dtype=torch.int32, device=dev)
                check(lib.b2me_block_rows(ptr(lv.coords), lv.V, key._ts, ptr(in2out1), ptr(in2out2), lv2.V, ptr(brows),
                                          stream()), "block_rows")
                check(lib.b2me_kernel_map_k3_blocks(ptr(lv.coords), lv.V, key._ts, ptr(lv2.table), lv2.table.numel(),
                                                    ptr(brows), ptr(nbr), ptr(row_masks), ptr(counts), stream()),
                      "kernel_map_k3_blocks")
                _count(4)
            else:
                check(lib.b2me_kernel_map_k3(ptr(lv.coords), lv.V, key._ts, ptr(lv.table), lv.table.numel(), ptr(nbr),
                                             None, stream()), "kernel_map_k3")
                check(lib.b2me_row_masks(ptr(nbr), lv.V, 27, ptr(row_masks), ptr(counts), stream()), "row_masks")
                _count(3)
            lv.nbr_k3 = nbr[:lv.V]
            lv.masks_k3 = (row_masks[:lv.V], counts)
        return lv.nbr_k3

    def perm_k3(self, key):
        """(row order, tile masks) of the k3 kernel map for the tcgen05 convolution."""
        lv = self.levels[key]
        if lv.perm_k3 is None:
            nbr = self.kernel_map_k3(key)
            row_masks, counts = lv.masks_k3
            perm = (mask_sorted_perm(nbr, lv.V, 27, coords=lv.coords, ts=key._ts, row_masks=row_masks,
                                     offset_counts=counts)
                    if (_State.mask_sort and lv.V >= _State.mask_sort_min_rows) else None)
            lv.perm_k3 = (perm, tile_masks(nbr, perm, lv.V, 27, row_masks=row_masks))
        return lv.perm_k3

    def perm_stride(self, rec, which):
        """(row order, tile masks) of the k2 s2 ("down") / transposed ("up") kernel map of a stride record."""
        name = "perm_" + which
        if rec.get(name) is None:
            nbr = rec["nbr_" + which]
            perm = (mask_sorted_perm(nbr, nbr.shape[0], 8)
                    if (_State.mask_sort and nbr.shape[0] >= _State.mask_sort_min_rows) else None)
            rec[name] = (perm, tile_masks(nbr, perm, nbr.shape[0], 8))
        return rec[name]

    # -- K2
    def stride_down(self, key):
        """coordinate map of a k=2, s=2 convolution on `key`; returns (coarse_key, down-record)."""
        lv = self.levels[key]
        if lv.down is None:
            dev = lv.coords.device
            V_in = lv.V
            cap = max(V_in, 1)
            out_coords = torch.empty((cap, 4), dtype=torch.int32, device=dev)
            in2out = torch.empty((cap,), dtype=torch.int32, device=dev)
            koff = torch.empty((cap,), dtype=torch.uint8, device=dev)
            counts = torch.zeros((4,), dtype=torch.int32, device=dev)
            table = torch.empty((lib.b2me_table_bytes(V_in),), dtype=torch.uint8, device=dev)
            ws = torch.empty((lib.b2me_unique_workspace_bytes(V_in, 0),), dtype=torch.uint8, device=dev)
            ts_out = key._ts * 2
            check(lib.b2me_stride_map(ptr(lv.coords), V_in, ts_out, ptr(out_coords), ptr(in2out), ptr(koff),
                                      ptr(counts), ptr(table), table.numel(), ptr(ws), ws.numel(), stream()),
                  "stride_map")
            _count(8)
            V_out, err = counts[:2].tolist()
            if err:
                raise _lib.B2MEError("stride map: coordinate outside key range")
            nbr_down = torch.empty((max(V_out, 1), 8), dtype=torch.int32, device=dev)
            nbr_up = torch.empty((cap, 8), dtype=torch.int32, device=dev)
            check(lib.b2me_stride_kernel_maps(ptr(in2out), ptr(koff), V_in, V_out, ptr(nbr_down), ptr(nbr_up),
                                              stream()), "stride_kernel_maps")
            _count(3)
            ckey = CoordinateMapKey(ts_out, key._tag)
            if ckey not in self.levels:
                self.levels[ckey] = _Level(out_coords[:V_out], table, V_out)
            lv.down = dict(in2out=in2out[:V_in], koff=koff[:V_in], nbr_down=nbr_down[:V_out], nbr_up=nbr_up[:V_in],
                           coarse_key=ckey)
        return lv.down["coarse_key"], lv.down

    def stride_up(self, key):
        """the cached fine map a transposed k=2, s=2 convolution lands on (model/backbone/minkunet.py:87-109)."""
        ts = key._ts
        if ts % 2:
            raise NotImplementedError("transposed convolution below tensor stride 1")
        fkey = CoordinateMapKey(ts // 2, key._tag)
        lv = self.levels.get(fkey)
        if lv is None or lv.down is None or lv.down["coarse_key"] != key:
            raise NotImplementedError("MinkowskiConvolutionTranspose onto a coordinate map that was not produced by "
                                      "a preceding stride-2 convolution (generative transposed conv is out of scope)")
        return fkey, lv.down

    def batch_size(self, key):
        lv = self.levels[key]
        if lv.batch_size is None:
            if self.number_of_batches is not None:
                lv.batch_size = self.number_of_batches
            else:
                lv.batch_size = int(lv.coords[:, 0].max().item()) + 1 if lv.V else 0
        return lv.batch_size


# ------------------------------------------------------------------------------------------------ pending ops
@dataclass
class _Pending:
    kind: str                      # "conv" | "elt"
    src: List["SparseTensor"]      # 1 or 2 sources (2 = lazy cat)
    module: object = None          # owner of the weights (conv / linear)
    nbr: Optional[torch.Tensor] = None
    perm: object = None            # callable returning (row order, tile masks) of `nbr` (built on first use)
    K: int = 1
    V_out: int = 0
    Cout: int = 0
    scale: Optional[torch.Tensor] = None
    shift: Optional[torch.Tensor] = None
    residual: Optional["SparseTensor"] = None
    act: int = _lib.ACT_NONE
    slope: float = 0.0
    stage: int = 0                 # 0 raw, 1 affine applied, 2 residual added, 3 activation applied
    extra: dict = dc_field(default_factory=dict)

    def clone(self):
        return _Pending(self.kind, list(self.src), self.module, self.nbr, self.perm, self.K, self.V_out, self.Cout,
                        self.scale, self.shift, self.residual, self.act, self.slope, self.stage, dict(self.extra))


def _as_dtype(t, dtype):
    """feature tensor -> dtype through the library's conversion kernel."""
    if t.dtype == dtype:
        return t
    out = torch.empty(t.shape, dtype=dtype, device=t.device)
    check(lib.b2me_convert(ptr(t), dtype_code(t.dtype), ptr(out), dtype_code(dtype), t.numel(), stream()), "convert")
    _count(1)
    return out


class SparseTensor:
    """Drop-in for ME.SparseTensor (the subset SURVEY.md §8b lists)."""

    def __init__(self, features=None, coordinates=None, tensor_stride=1, coordinate_map_key=None,
                 coordinate_manager=None, quantization_mode=SparseTensorQuantizationMode.RANDOM_SUBSAMPLE,
                 allow_duplicate_coordinates=False, minkowski_algorithm=MinkowskiAlgorithm.DEFAULT,
                 requires_grad=None, device=None, _pending=None, _cat=None):
        self._pending = _pending
        self._cat = _cat
        self._F = None
        self.quantization_mode = quantization_mode
        self.inverse_mapping = None
        if coordinate_map_key is not None:
            assert coordinate_manager is not None
            self.coordinate_map_key = coordinate_map_key
            self.coordinate_manager = coordinate_manager
            if features is not None:
                self._F = features
            return
        if features is None or coordinates is None:
            raise ValueError("SparseTensor needs features and coordinates (or a coordinate_map_key)")
        if requires_grad:
            raise NotImplementedError("this MinkowskiEngine build is inference-only (no backward kernels)")
        if isinstance(tensor_stride, (list, tuple)):
            tensor_stride = tensor_stride[0]
        if tensor_stride != 1:
            raise NotImplementedError("SparseTensor construction at tensor_stride != 1")
        dev = torch.device(device) if device is not None else features.device
        if dev.type != "cuda":
            raise _lib.B2MEError("SparseTensor must be created on a CUDA device (no CPU fallback)")
        coordinates = coordinates.to(dev)
        features = features.to(dev)
        if coordinates.dtype.is_floating_point:
            coordinates = torch.floor(coordinates).to(torch.int32)
        mgr = coordinate_manager or CoordinateManager(D=coordinates.shape[1] - 1, device=dev)
        mode = 1 if quantization_mode == SparseTensorQuantizationMode.UNWEIGHTED_AVERAGE else 0
        if quantization_mode not in (SparseTensorQuantizationMode.RANDOM_SUBSAMPLE,
                                     SparseTensorQuantizationMode.UNWEIGHTED_AVERAGE):
            raise NotImplementedError(f"quantization mode {quantization_mode}")
        key, vf, inverse, first_idx = mgr.insert_coordinates(coordinates.to(torch.int32), features, mode)
        self.coordinate_map_key = key
        self.coordinate_manager = mgr
        self._F = vf
        self.inverse_mapping = inverse
        self.unique_index = first_idx

    # ---- lazy evaluation
    def _materialize(self):
        if self._F is not None:
            return self._F
        if self._cat is not None:
            self._F = torch.cat([s._materialize() for s in self._cat], dim=1)
            self._cat = None
            return self._F
        p = self._pending
        assert p is not None
        from . import ops
        self._F = ops.run_pending(p)
        self._pending = None
        return self._F

    @property
    def F(self):
        return self._materialize()

    features = F

    @property
    def C(self):
        return self.coordinate_manager.coordinates(self.coordinate_map_key)

    coordinates = C

    @property
    def tensor_stride(self):
        return self.coordinate_map_key.get_tensor_stride()

    @property
    def D(self):
        return 3

    @property
    def device(self):
        return self.C.device

    @property
    def dtype(self):
        return self.F.dtype

    @property
    def num_channels(self):
        if self._F is not None:
            return self._F.shape[1]
        if self._cat is not None:
            return sum(s.num_channels for s in self._cat)
        return self._pending.Cout

    @property
    def num_rows(self):
        return self.coordinate_manager.level(self.coordinate_map_key).V

    @property
    def shape(self):
        return torch.Size([self.num_rows, self.num_channels])

    def size(self, dim=None):
        return self.shape if dim is None else self.shape[dim]

    def __len__(self):
        return self.num_rows

    def __repr__(self):
        return (f"SparseTensor(rows={self.num_rows}, channels={self.num_channels}, "
                f"tensor_stride={self.tensor_stride}, lazy={self._F is None})")

    def _same_map(self, other):
        return (self.coordinate_manager is other.coordinate_manager
                and self.coordinate_map_key == other.coordinate_map_key)

    def _child(self, **kw):
        return SparseTensor(coordinate_map_key=self.coordinate_map_key, coordinate_manager=self.coordinate_manager,
                            **kw)

    # out += residual (BasicBlock): folded into the producing kernel's epilogue when possible
    def __add__(self, other):
        if not isinstance(other, SparseTensor):
            return self._child(features=self.F + other)
        if not self._same_map(other):
            raise NotImplementedError("adding SparseTensors on different coordinate maps")
        p = self._pending
        if self._F is None and p is not None and p.stage <= 1 and p.residual is None:
            q = p.clone()
            q.residual = other
            q.stage = 2
            return self._child(_pending=q)
        q = _Pending("elt", [self], V_out=self.num_rows, Cout=self.num_channels, residual=other, stage=2)
        return self._child(_pending=q)

    __iadd__ = __add__

    def slice(self, tensor_field):
        """features of every point of the field = features of its voxel (app/inference_engine.py:417)."""
        if self.coordinate_map_key._ts != 1:
            raise ValueError("slice needs a tensor-stride-1 SparseTensor")
        inverse = tensor_field.inverse_mapping
        if inverse is None:
            raise ValueError("field has no inverse mapping; call field.sparse() first")
        F = self.F
        N = inverse.shape[0]
        out = torch.empty((N, F.shape[1]), dtype=F.dtype, device=F.device)
        check(lib.b2me_gather_rows(ptr(F), dtype_code(F.dtype), F.shape[1], ptr(inverse), N, ptr(out), stream()),
              "gather_rows")
        _count(1)
        return TensorField(features=out, _inverse=inverse, _manager=self.coordinate_manager,
                           _coordinates=tensor_field._coordinates)

    def _batch_rows(self):
        C = self.C
        B = self.coordinate_manager.batch_size(self.coordinate_map_key)
        return [torch.nonzero(C[:, 0] == b).flatten() for b in range(B)]

    @property
    def decomposed_coordinates(self):
        C = self.C
        return [C[r, 1:] for r in self._batch_rows()]

    @property
    def decomposed_features(self):
        F = self.F
        return [F[r] for r in self._batch_rows()]

    @property
    def decomposed_coordinates_and_features(self):
        return self.decomposed_coordinates, self.decomposed_features

    def dense(self, *a, **k):
        raise NotImplementedError("SparseTensor.dense is outside the hot-path subset")


class TensorField:
    """Drop-in for ME.TensorField: raw (un-quantised) points; .sparse() voxelises them (K1)."""

    def __init__(self, features=None, coordinates=None, tensor_stride=1, coordinate_field_map_key=None,
                 coordinate_manager=None, quantization_mode=SparseTensorQuantizationMode.UNWEIGHTED_AVERAGE,
                 allow_duplicate_coordinates=False, minkowski_algorithm=MinkowskiAlgorithm.DEFAULT,
                 requires_grad=None, device=None, _inverse=None, _manager=None, _coordinates=None):
        self.quantization_mode = quantization_mode
        self.inverse_mapping = _inverse
        self.coordinate_manager = _manager or coordinate_manager
        if _inverse is not None:  # produced by SparseTensor.slice
            self._features = features
            self._coordinates = _coordinates
            return
        if features is None or coordinates is None:
            raise ValueError("TensorField needs features and coordinates")
        dev = torch.device(device) if device is not None else features.device
        if dev.type != "cuda":
            raise _lib.B2MEError("TensorField must be created on a CUDA device (no CPU fallback)")
        self._nb = None
        if not coordinates.is_cuda and coordinates.shape[0]:
            self._nb = int(coordinates[:, 0].max().item()) + 1  # free on the host, saves a device sync later
        self._features = features.to(dev, non_blocking=True)
        self._coordinates = coordinates.to(dev, non_blocking=True)

    @property
    def F(self):
        return self._features

    features = F

    @property
    def C(self):
        return self._coordinates

    coordinates = C

    @property
    def device(self):
        return self._features.device

    def sparse(self, tensor_stride=1, coordinate_map_key=None, quantization_mode=None):
        if quantization_mode is None:
            quantization_mode = self.quantization_mode
        if quantization_mode == SparseTensorQuantizationMode.UNWEIGHTED_AVERAGE:
            mode = 1
        elif quantization_mode == SparseTensorQuantizationMode.RANDOM_SUBSAMPLE:
            mode = 0
        else:
            raise NotImplementedError(f"quantization mode {quantization_mode}")
        mgr = CoordinateManager(D=self._coordinates.shape[1] - 1, device=self._features.device)
        mgr.number_of_batches = getattr(self, "_nb", None)
        key, vf, inverse, first_idx = mgr.insert_coordinates(self._coordinates, self._features, mode)
        self.inverse_mapping = inverse
        self.coordinate_manager = mgr
        mgr.field_inverse = inverse
        st = SparseTensor(features=vf, coordinate_map_key=key, coordinate_manager=mgr)
        st.inverse_mapping = inverse
        st.unique_index = first_idx
        return st


def cat(*tensors):
    """ME.cat: channel concatenation of tensors on the same coordinate map (minkunet.py:156-180).
    Lazy: a following convolution reads both sources directly."""
    if len(tensors) == 1 and isinstance(tensors[0], (list, tuple)):
        tensors = tuple(tensors[0])
    first = tensors[0]
    for t in tensors[1:]:
        if not first._same_map(t):
            raise ValueError("ME.cat needs tensors on the same coordinate map")
    if len(tensors) == 1:
        return first
    if len(tensors) == 2 and all(t._cat is None for t in tensors):
        return first._child(_cat=list(tensors))
    return first._child(features=torch.cat([t.F for t in tensors], dim=1))
