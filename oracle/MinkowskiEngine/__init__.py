"""CPU oracle of the MinkowskiEngine subset (TEST INFRASTRUCTURE, see oracle/__init__.py).

Eager PyTorch-CPU / NumPy restatement of the semantics SURVEY.md §8(a) rows a1-a14 state for
MinkowskiEngine 0.5.4 as called from model/backbone/minkunet.py:52-187, model/backbone/resnet.py:86-127,
model/robotnet_{segmentation,vote,encode}.py and app/inference_engine.py:405-417:
  * voxel rows in FIRST-OCCURRENCE order of the input points; UNWEIGHTED_AVERAGE = per-voxel mean
  * stride-2 maps = unique(floor(c / 2ts) * 2ts), first-occurrence order
  * kernel offsets: odd kernels centred, even kernels start at 0, x fastest in the kernel index
  * transposed k2 s2 convolution lands on the cached fine map with the forward map's offsets
"""
__version__ = "0.5.4+oracle"

import math
from enum import Enum

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as TF

FIXED_ONE = 4294967296.0  # 2^32, the fixed-point scale of the voxel mean


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


# ------------------------------------------------------------------------------------------------ coordinates
_BIAS = 1 << 17


def pack_keys(c):
    """[V,4] int (b,x,y,z) -> int64 key, same field layout as the CUDA hash (b:10|x:18|y:18|z:18)."""
    c = np.asarray(c, dtype=np.int64)
    return (c[:, 0] << 54) | ((c[:, 1] + _BIAS) << 36) | ((c[:, 2] + _BIAS) << 18) | (c[:, 3] + _BIAS)


def unique_first_occurrence(q):
    """rows of q [N,4] int -> (unique rows in first-occurrence order, inverse [N], first index [V])."""
    q = np.asarray(q, dtype=np.int64)
    if len(q) == 0:
        return q.reshape(0, 4).astype(np.int32), np.zeros(0, np.int64), np.zeros(0, np.int64)
    keys = pack_keys(q)
    _, first, inv = np.unique(keys, return_index=True, return_inverse=True)
    order = np.argsort(first, kind="stable")       # sorted-unique id -> position in first-occurrence order
    rank = np.empty_like(order)
    rank[order] = np.arange(len(order))
    return q[first[order]].astype(np.int32), rank[inv.reshape(-1)], first[order]


def voxel_mean_fixed_point(feats, inverse, V):
    """per-voxel mean exactly as K1 computes it: sum of round(x * 2^32) in int64, / 2^32 / count in double."""
    f = np.asarray(feats, dtype=np.float32)
    fx = np.rint(f.astype(np.float64) * FIXED_ONE).astype(np.int64)
    sums = np.zeros((V, f.shape[1]), dtype=np.int64)
    np.add.at(sums, inverse, fx)
    cnt = np.bincount(inverse, minlength=V).astype(np.float64)
    return ((sums.astype(np.float64) / FIXED_ONE) / cnt[:, None]).astype(np.float32)


class CoordinateMapKey:
    def __init__(self, ts, tag=""):
        self._ts, self._tag = int(ts), tag

    def get_tensor_stride(self):
        return [self._ts] * 3

    def __eq__(self, o):
        return isinstance(o, CoordinateMapKey) and (self._ts, self._tag) == (o._ts, o._tag)

    def __hash__(self):
        return hash((self._ts, self._tag))


class _Level:
    def __init__(self, coords):
        self.coords = np.ascontiguousarray(coords, dtype=np.int32)
        keys = pack_keys(self.coords)
        self.sort = np.argsort(keys, kind="stable")
        self.skeys = keys[self.sort]
        self.nbr_k3 = None
        self.down = None

    @property
    def V(self):
        return len(self.coords)

    def lookup(self, c):
        """rows of the coordinates c [M,4] in this map, -1 where absent."""
        k = pack_keys(c)
        if len(self.skeys) == 0:
            return np.full(len(k), -1, np.int64)
        pos = np.searchsorted(self.skeys, k)
        pos = np.minimum(pos, len(self.skeys) - 1)
        hit = self.skeys[pos] == k
        return np.where(hit, self.sort[pos], -1)


class CoordinateManager:
    def __init__(self):
        self.levels = {}

    def kernel_map_k3(self, key):
        lv = self.levels[key]
        if lv.nbr_k3 is None:
            ts = key._ts
            nbr = np.empty((lv.V, 27), dtype=np.int64)
            for k in range(27):
                dx, dy, dz = k % 3 - 1, (k // 3) % 3 - 1, k // 9 - 1   # x fastest
                c = lv.coords.astype(np.int64).copy()
                c[:, 1] += dx * ts
                c[:, 2] += dy * ts
                c[:, 3] += dz * ts
                nbr[:, k] = lv.lookup(c)
            lv.nbr_k3 = nbr
        return lv.nbr_k3

    def stride_down(self, key):
        lv = self.levels[key]
        if lv.down is None:
            ts_in, ts_out = key._ts, key._ts * 2
            c = lv.coords.astype(np.int64)
            parent = c.copy()
            parent[:, 1:] = np.floor_divide(c[:, 1:], ts_out) * ts_out
            pc, in2out, _ = unique_first_occurrence(parent)
            d = (c[:, 1:] - parent[:, 1:]) // ts_in
            koff = d[:, 0] + 2 * d[:, 1] + 4 * d[:, 2]
            nbr_down = np.full((len(pc), 8), -1, dtype=np.int64)
            nbr_down[in2out, koff] = np.arange(len(c))
            nbr_up = np.full((len(c), 8), -1, dtype=np.int64)
            nbr_up[np.arange(len(c)), koff] = in2out
            ckey = CoordinateMapKey(ts_out, key._tag)
            if ckey not in self.levels:
                self.levels[ckey] = _Level(pc)
            lv.down = dict(in2out=in2out, koff=koff, nbr_down=nbr_down, nbr_up=nbr_up, coarse_key=ckey)
        return lv.down["coarse_key"], lv.down

    def stride_up(self, key):
        fkey = CoordinateMapKey(key._ts // 2, key._tag)
        lv = self.levels.get(fkey)
        if lv is None or lv.down is None:
            raise NotImplementedError("transposed convolution onto an uncached map")
        return fkey, lv.down


# ------------------------------------------------------------------------------------------------ tensors
class SparseTensor:
    def __init__(self, features=None, coordinates=None, tensor_stride=1, coordinate_map_key=None,
                 coordinate_manager=None, quantization_mode=SparseTensorQuantizationMode.RANDOM_SUBSAMPLE,
                 allow_duplicate_coordinates=False, minkowski_algorithm=MinkowskiAlgorithm.DEFAULT,
                 requires_grad=None, device=None):
        self.inverse_mapping = None
        if coordinate_map_key is not None:
            self.F = features
            self.coordinate_map_key = coordinate_map_key
            self.coordinate_manager = coordinate_manager
            return
        c = coordinates.detach().cpu()
        q = torch.floor(c).to(torch.int32).numpy() if c.dtype.is_floating_point else c.to(torch.int32).numpy()
        uc, inv, first = unique_first_occurrence(q)
        f = features.detach().cpu().float()
        if quantization_mode == SparseTensorQuantizationMode.UNWEIGHTED_AVERAGE:
            vf = torch.from_numpy(voxel_mean_fixed_point(f.numpy(), inv, len(uc)))
        else:
            vf = f[torch.from_numpy(first)]
        mgr = CoordinateManager()
        key = CoordinateMapKey(1)
        mgr.levels[key] = _Level(uc)
        self.F = vf
        self.coordinate_map_key = key
        self.coordinate_manager = mgr
        self.inverse_mapping = torch.from_numpy(inv)
        self.unique_index = torch.from_numpy(first)

    @property
    def features(self):
        return self.F

    @property
    def C(self):
        return torch.from_numpy(self.coordinate_manager.levels[self.coordinate_map_key].coords)

    coordinates = C

    @property
    def tensor_stride(self):
        return self.coordinate_map_key.get_tensor_stride()

    @property
    def D(self):
        return 3

    @property
    def device(self):
        return self.F.device

    @property
    def dtype(self):
        return self.F.dtype

    @property
    def shape(self):
        return self.F.shape

    def _child(self, F):
        return SparseTensor(F, coordinate_map_key=self.coordinate_map_key,
                            coordinate_manager=self.coordinate_manager)

    def __add__(self, other):
        if isinstance(other, SparseTensor):
            assert other.coordinate_map_key == self.coordinate_map_key
            return self._child(self.F + other.F)
        return self._child(self.F + other)

    __iadd__ = __add__

    def slice(self, field):
        assert self.coordinate_map_key._ts == 1
        return TensorField(features=self.F[field.inverse_mapping], _inverse=field.inverse_mapping)

    @property
    def decomposed_coordinates(self):
        C = self.C
        B = int(C[:, 0].max()) + 1 if len(C) else 0
        return [C[C[:, 0] == b, 1:] for b in range(B)]

    @property
    def decomposed_features(self):
        C = self.C
        B = int(C[:, 0].max()) + 1 if len(C) else 0
        return [self.F[C[:, 0] == b] for b in range(B)]


class TensorField:
    def __init__(self, features=None, coordinates=None, tensor_stride=1, coordinate_field_map_key=None,
                 coordinate_manager=None, quantization_mode=SparseTensorQuantizationMode.UNWEIGHTED_AVERAGE,
                 allow_duplicate_coordinates=False, minkowski_algorithm=MinkowskiAlgorithm.DEFAULT,
                 requires_grad=None, device=None, _inverse=None):
        self.F = features.detach().cpu() if features is not None else None
        self._coordinates = coordinates.detach().cpu() if coordinates is not None else None
        self.quantization_mode = quantization_mode
        self.inverse_mapping = _inverse

    @property
    def features(self):
        return self.F

    @property
    def C(self):
        return self._coordinates

    def sparse(self, tensor_stride=1, coordinate_map_key=None, quantization_mode=None):
        st = SparseTensor(self.F, self._coordinates, quantization_mode=quantization_mode or self.quantization_mode)
        self.inverse_mapping = st.inverse_mapping
        return st


def cat(*tensors):
    if len(tensors) == 1 and isinstance(tensors[0], (list, tuple)):
        tensors = tuple(tensors[0])
    for t in tensors[1:]:
        assert t.coordinate_map_key == tensors[0].coordinate_map_key, "ME.cat: different coordinate maps"
    return tensors[0]._child(torch.cat([t.F for t in tensors], dim=1))


# ------------------------------------------------------------------------------------------------ modules
def _scalar(v):
    return int(v[0]) if isinstance(v, (list, tuple)) else int(v)


def sparse_conv(F_in, W, nbr, V_out):
    """out[o] = sum_k in[nbr[o,k]] @ W[k]  (gather -> matmul -> index_add), fp32."""
    out = torch.zeros((V_out, W.shape[2]), dtype=torch.float32)
    if nbr is None:
        return F_in @ W[0]
    for k in range(W.shape[0]):
        col = nbr[:, k]
        rows = np.nonzero(col >= 0)[0]
        if len(rows) == 0:
            continue
        out.index_add_(0, torch.from_numpy(rows), F_in[torch.from_numpy(col[rows])] @ W[k])
    return out


class _ConvBase(nn.Module):
    is_transpose = False

    def __init__(self, in_channels, out_channels, kernel_size=-1, stride=1, dilation=1, bias=False,
                 kernel_generator=None, expand_coordinates=False, convolution_mode=None, dimension=None):
        super().__init__()
        assert dimension == 3
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel_size, self.stride, self.dilation = _scalar(kernel_size), _scalar(stride), _scalar(dilation)
        self.kernel_volume = self.kernel_size ** 3
        shape = (self.kernel_volume, in_channels, out_channels) if self.kernel_volume > 1 else (in_channels,
                                                                                                out_channels)
        self.kernel = nn.Parameter(torch.empty(*shape))
        self.bias = nn.Parameter(torch.empty(1, out_channels)) if bias else None
        n = (out_channels if self.is_transpose else in_channels) * self.kernel_volume
        stdv = 1.0 / math.sqrt(n)
        with torch.no_grad():
            self.kernel.uniform_(-stdv, stdv)
            if self.bias is not None:
                self.bias.uniform_(-stdv, stdv)

    def forward(self, x):
        mgr, key = x.coordinate_manager, x.coordinate_map_key
        ks, st = self.kernel_size, self.stride
        assert self.dilation == 1
        if not self.is_transpose:
            if ks == 1 and st == 1:
                out_key, nbr = key, None
            elif ks == 3 and st == 1:
                out_key, nbr = key, mgr.kernel_map_k3(key)
            elif ks == 2 and st == 2:
                out_key, rec = mgr.stride_down(key)
                nbr = rec["nbr_down"]
            else:
                raise NotImplementedError((ks, st))
        else:
            assert ks == 2 and st == 2
            out_key, rec = mgr.stride_up(key)
            nbr = rec["nbr_up"]
        W = self.kernel.detach().float()
        if W.dim() == 2:
            W = W.unsqueeze(0)
        out = sparse_conv(x.F.float(), W, nbr, mgr.levels[out_key].V)
        if self.bias is not None:
            out = out + self.bias.detach().float()
        return SparseTensor(out, coordinate_map_key=out_key, coordinate_manager=mgr)


class MinkowskiConvolution(_ConvBase):
    is_transpose = False


class MinkowskiConvolutionTranspose(_ConvBase):
    is_transpose = True


class MinkowskiBatchNorm(nn.Module):
    def __init__(self, num_features, eps=1e-5, momentum=0.1, affine=True, track_running_stats=True):
        super().__init__()
        self.bn = nn.BatchNorm1d(num_features, eps=eps, momentum=momentum, affine=affine,
                                 track_running_stats=track_running_stats)

    def forward(self, x):
        return x._child(self.bn(x.F))


class MinkowskiReLU(nn.Module):
    def __init__(self, inplace=False):
        super().__init__()

    def forward(self, x):
        return x._child(TF.relu(x.F))


class MinkowskiLeakyReLU(nn.Module):
    def __init__(self, negative_slope=0.01, inplace=False):
        super().__init__()
        self.negative_slope = negative_slope

    def forward(self, x):
        return x._child(TF.leaky_relu(x.F, self.negative_slope))


class MinkowskiSigmoid(nn.Module):
    def forward(self, x):
        return x._child(torch.sigmoid(x.F))


class MinkowskiLinear(nn.Module):
    def __init__(self, in_features, out_features, bias=True):
        super().__init__()
        self.linear = nn.Linear(in_features, out_features, bias=bias)

    def forward(self, x):
        return x._child(self.linear(x.F))


class _GlobalPool(nn.Module):
    mode = "avg"

    def __init__(self, mode=None):
        super().__init__()

    def forward(self, x):
        C = x.C
        B = int(C[:, 0].max()) + 1 if len(C) else 0
        rows = []
        for b in range(B):
            f = x.F[C[:, 0] == b]
            rows.append(f.mean(0) if self.mode == "avg" else f.max(0).values)
        out = torch.stack(rows) if rows else torch.zeros((0, x.F.shape[1]))
        mgr = x.coordinate_manager
        gkey = CoordinateMapKey(0, "global")
        gc = np.zeros((B, 4), np.int32)
        gc[:, 0] = np.arange(B)
        mgr.levels[gkey] = _Level(gc)
        return SparseTensor(out, coordinate_map_key=gkey, coordinate_manager=mgr)


class MinkowskiGlobalAvgPooling(_GlobalPool):
    mode = "avg"


class MinkowskiGlobalMaxPooling(_GlobalPool):
    mode = "max"


def _placeholder(name):
    class _P(nn.Module):
        def __init__(self, *a, **k):
            super().__init__()

        def forward(self, *a, **k):
            raise NotImplementedError(name)

    _P.__name__ = name
    return _P


for _n in ("MinkowskiInstanceNorm", "MinkowskiMaxPooling", "MinkowskiAvgPooling", "MinkowskiDropout", "MinkowskiGELU",
           "MinkowskiSinusoidal", "MinkowskiToSparseTensor", "MinkowskiSumPooling"):
    globals()[_n] = _placeholder(_n)

from . import utils, modules, MinkowskiOps  # noqa: E402,F401
