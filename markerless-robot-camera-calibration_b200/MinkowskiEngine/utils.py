"""ME.utils subset: batched_coordinates, sparse_collate, sparse_quantize, kaiming_normal_
(data/alivev2.py:289-296,358-414; model/backbone/resnet.py:86-90)."""
import math

import numpy as np
import torch

from . import _lib
from ._lib import lib, check, ptr, stream


def batched_coordinates(coords, dtype=torch.int32, device=None):
    """list of [N_i, D] -> [sum N_i, 1+D] with the batch index in column 0; integer dtypes floor first."""
    if dtype not in (torch.int32, torch.float32):
        raise ValueError("dtype must be torch.int32 or torch.float32")
    out = []
    for b, c in enumerate(coords):
        if isinstance(c, np.ndarray):
            c = torch.from_numpy(c)
        c = torch.as_tensor(c)
        if dtype == torch.int32:
            c = torch.floor(c).to(torch.int32) if c.dtype.is_floating_point else c.to(torch.int32)
        else:
            c = c.to(torch.float32)
        bcol = torch.full((c.shape[0], 1), b, dtype=dtype, device=c.device)
        out.append(torch.cat((bcol, c), dim=1))
    res = torch.cat(out, dim=0) if out else torch.zeros((0, 4), dtype=dtype)
    return res.to(device) if device is not None else res


def sparse_collate(coords, feats, labels=None, dtype=torch.int32, device=None):
    def cat_list(lst):
        ts = [torch.from_numpy(x) if isinstance(x, np.ndarray) else torch.as_tensor(x) for x in lst]
        r = torch.cat(ts, dim=0)
        return r.to(device) if device is not None else r

    bcoords = batched_coordinates(coords, dtype=dtype, device=device)
    bfeats = cat_list(feats)
    if labels is None:
        return bcoords, bfeats
    return bcoords, bfeats, cat_list(labels)


def sparse_quantize(coordinates, features=None, labels=None, ignore_label=-100, return_index=False,
                    return_inverse=False, return_maps_only=False, quantization_size=None, device="cuda"):
    """Voxelise on the GPU (K1, first point of each voxel wins; rows in first-occurrence order).
    Returns in the container type of the input (numpy in -> numpy out), like ME."""
    is_np = isinstance(coordinates, np.ndarray)
    c = torch.from_numpy(coordinates) if is_np else coordinates
    if quantization_size is not None:
        # this entry point DIVIDES (data/alivev2.py:290-296), unlike the TensorField path which multiplies
        c = c / quantization_size if c.dtype.is_floating_point else c.double() / quantization_size
    q = torch.floor(c).to(torch.int32) if c.dtype.is_floating_point else c.to(torch.int32)
    dev = torch.device(device if str(device) != "cpu" else "cuda")
    if not torch.cuda.is_available():
        raise _lib.B2MEError("sparse_quantize runs on the GPU; no CUDA device is visible (no CPU fallback)")
    N, D = q.shape
    q4 = torch.zeros((N, 4), dtype=torch.int32)
    q4[:, 4 - D:] = q
    q4 = q4.to(dev)
    cap = max(N, 1)
    out_coords = torch.empty((cap, 4), dtype=torch.int32, device=dev)
    inverse = torch.empty((cap,), dtype=torch.int32, device=dev)
    first_idx = torch.empty((cap,), dtype=torch.int32, device=dev)
    counts = torch.zeros((4,), dtype=torch.int32, device=dev)
    table = torch.empty((lib.b2me_table_bytes(N),), dtype=torch.uint8, device=dev)
    ws = torch.empty((lib.b2me_unique_workspace_bytes(N, 0),), dtype=torch.uint8, device=dev)
    check(lib.b2me_quantize_unique(None, ptr(q4), N, None, 0, 0, ptr(out_coords), None, ptr(inverse), ptr(first_idx),
                                   ptr(counts), ptr(table), table.numel(), ptr(ws), ws.numel(), stream()),
          "quantize_unique")
    V, err = counts[:2].tolist()
    if err:
        raise _lib.B2MEError("sparse_quantize: coordinate outside the key range")
    idx = first_idx[:V].long()
    inv = inverse[:N].long()
    vlabels = None
    if labels is not None:
        lab = (torch.from_numpy(labels) if isinstance(labels, np.ndarray) else labels).to(dev).to(torch.int32)
        vl = torch.empty((max(V, 1),), dtype=torch.int32, device=dev)
        check(lib.b2me_quantize_labels(ptr(lab.contiguous()), ptr(inverse), ptr(first_idx), N, V, int(ignore_label),
                                       ptr(vl), stream()), "quantize_labels")
        vlabels = vl[:V]

    def conv(t):
        t = t.cpu()
        return t.numpy() if is_np else t

    if return_maps_only:
        if return_inverse:
            return conv(idx), conv(inv)
        return conv(idx)
    res = [conv(out_coords[:V, 4 - D:])]
    if features is not None:
        f = torch.from_numpy(features) if isinstance(features, np.ndarray) else features
        fi = f[idx.cpu()]
        res.append(fi.numpy() if isinstance(features, np.ndarray) else fi)
    if vlabels is not None:
        vl = vlabels.cpu()
        res.append(vl.numpy().astype(labels.dtype) if isinstance(labels, np.ndarray) else vl.to(labels.dtype))
    if return_index:
        res.append(conv(idx))
    if return_inverse:
        res.append(conv(inv))
    return res[0] if len(res) == 1 else tuple(res)


def _fans(tensor):
    if tensor.dim() < 2:
        raise ValueError("fan in/out needs at least 2 dimensions")
    if tensor.dim() == 2:
        return tensor.size(1), tensor.size(0)
    k = tensor.size(0)
    return tensor.size(1) * k, tensor.size(2) * k


def kaiming_normal_(tensor, a=0, mode="fan_in", nonlinearity="leaky_relu"):
    """He-normal init with the fan of a [K, Cin, Cout] sparse kernel (model/backbone/resnet.py:86-90)."""
    fan_in, fan_out = _fans(tensor)
    fan = fan_in if mode == "fan_in" else fan_out
    gain = torch.nn.init.calculate_gain(nonlinearity, a)
    std = gain / math.sqrt(fan)
    with torch.no_grad():
        return tensor.normal_(0, std)
