"""Oracle ME.utils."""
import math

import numpy as np
import torch

from . import unique_first_occurrence


def batched_coordinates(coords, dtype=torch.int32, device=None):
    out = []
    for b, c in enumerate(coords):
        c = torch.as_tensor(np.asarray(c)) if not torch.is_tensor(c) else c
        if dtype == torch.int32:
            c = torch.floor(c).to(torch.int32) if c.dtype.is_floating_point else c.to(torch.int32)
        else:
            c = c.to(torch.float32)
        out.append(torch.cat((torch.full((len(c), 1), b, dtype=dtype), c), dim=1))
    return torch.cat(out, dim=0)


def sparse_collate(coords, feats, labels=None, dtype=torch.int32, device=None):
    tl = lambda lst: torch.cat([torch.as_tensor(np.asarray(x)) if not torch.is_tensor(x) else x for x in lst], 0)
    if labels is None:
        return batched_coordinates(coords, dtype), tl(feats)
    return batched_coordinates(coords, dtype), tl(feats), tl(labels)


def sparse_quantize(coordinates, features=None, labels=None, ignore_label=-100, return_index=False,
                    return_inverse=False, return_maps_only=False, quantization_size=None, device="cpu"):
    is_np = isinstance(coordinates, np.ndarray)
    c = coordinates if is_np else coordinates.numpy()
    if quantization_size is not None:
        c = c / quantization_size
    q = np.floor(c).astype(np.int32)
    N, D = q.shape
    q4 = np.zeros((N, 4), np.int32)
    q4[:, 4 - D:] = q
    uc, inv, first = unique_first_occurrence(q4)
    conv = (lambda a: a) if is_np else torch.from_numpy
    if return_maps_only:
        return (conv(first), conv(inv)) if return_inverse else conv(first)
    res = [conv(np.ascontiguousarray(uc[:, 4 - D:]))]
    if features is not None:
        res.append(features[first] if isinstance(features, np.ndarray) else features[torch.from_numpy(first)])
    if labels is not None:
        lab = labels if isinstance(labels, np.ndarray) else labels.numpy()
        vl = lab[first].copy()
        disagree = lab != vl[inv]
        vl[np.unique(inv[disagree])] = ignore_label
        res.append(vl if isinstance(labels, np.ndarray) else torch.from_numpy(vl))
    if return_index:
        res.append(conv(first))
    if return_inverse:
        res.append(conv(inv))
    return res[0] if len(res) == 1 else tuple(res)


def kaiming_normal_(tensor, a=0, mode="fan_in", nonlinearity="leaky_relu"):
    if tensor.dim() == 2:
        fan_in, fan_out = tensor.size(1), tensor.size(0)
    else:
        fan_in, fan_out = tensor.size(1) * tensor.size(0), tensor.size(2) * tensor.size(0)
    std = torch.nn.init.calculate_gain(nonlinearity, a) / math.sqrt(fan_in if mode == "fan_in" else fan_out)
    with torch.no_grad():
        return tensor.normal_(0, std)
