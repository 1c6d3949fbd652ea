"""TEST INFRASTRUCTURE (oracle): NumPy restatement of the K3b row-order keys that markerless-robot-camera-calibration_b200/
csrc/coords.cu computes for the tcgen05 convolution (b2me_mask_sort_keys / b2me_mask_sort_keys2). The row order is an
internal optimisation of this framework (the reference's MinkowskiEngine has no such step: any order gives the same
convolution result), so this file pins the DEFINITION the kernels implement, not a reference behaviour. Only tests may
import it."""
import numpy as np

TWO_LEVEL_T = 10  # MS2_T in coords.cu


def reflect(m):
    """bit i of the result = bit i of m XOR the parity of the bits of m above i (reflected / Gray order)."""
    m = m.astype(np.uint32)
    p = m >> np.uint32(1)
    for s in (1, 2, 4, 8, 16):
        p ^= p >> np.uint32(s)
    return m ^ p


def _global_keys(nbr):
    """neighbour-occupancy masks with the bits reordered: the rarest offset of the map is the most significant bit
    (ties: lower offset index is the more significant)."""
    V, K = nbr.shape
    pres = nbr >= 0
    counts = pres.sum(0)
    rank = np.zeros(K, np.int64)
    for k in range(K):
        rank[k] = sum(1 for j in range(K) if counts[j] < counts[k] or (counts[j] == counts[k] and j < k))
    bitpos = K - 1 - rank
    key = np.zeros(V, np.uint32)
    for k in range(K):
        key |= pres[:, k].astype(np.uint32) << np.uint32(bitpos[k])
    return key


def keys_one_level(nbr):
    return reflect(_global_keys(nbr)).astype(np.int32)


def keys_two_level(nbr, T=TWO_LEVEL_T):
    V, K = nbr.shape
    if K <= T + 1:
        return keys_one_level(nbr)
    R = K - T
    key = _global_keys(nbr)
    seg = key >> np.uint32(R)
    low = np.zeros(V, np.uint32)
    for s in np.unique(seg):
        rows = np.flatnonzero(seg == s)
        sub = key[rows]
        tot = len(rows)
        c = np.array([int(((sub >> np.uint32(b)) & 1).sum()) for b in range(R)], dtype=np.int64)
        c[(c == 0) | (c == tot)] = 0xFFFFFFFF  # offsets that no row / every row of the segment has go last
        lo = np.zeros(tot, np.uint32)
        for b in range(R):
            rank = sum(1 for j in range(R) if c[j] < c[b] or (c[j] == c[b] and j > b))
            lo |= ((sub >> np.uint32(b)) & np.uint32(1)) << np.uint32(R - 1 - rank)
        low[rows] = lo
    return ((reflect(seg) << np.uint32(R)) | reflect(low)).astype(np.int32)


def passes(nbr, order, rows_per_tile=256):
    """(tile, offset) passes the convolution executes for this row order."""
    V, K = nbr.shape
    pres = nbr[order] >= 0
    pad = (-V) % rows_per_tile
    pres = np.concatenate((pres, np.zeros((pad, K), bool)))
    return int(pres.reshape(-1, rows_per_tile, K).any(1).sum())


def _spread10(v):
    v = v.astype(np.uint64) & np.uint64(0x3FF)
    out = np.zeros_like(v)
    for b in range(10):
        out |= ((v >> np.uint64(b)) & np.uint64(1)) << np.uint64(3 * b)
    return out


def keys_morton(nbr, coords, ts=1):
    """b2me_mask_sort_keys_morton: reflected mask key << 37 | frame (7 bits) << 30 | 30-bit Morton code of coord / ts
    (x fastest, 10 wrapped bits per axis), as a signed 64-bit key."""
    shift = int(ts).bit_length() - 1
    c = coords.astype(np.int64)
    m = (_spread10(c[:, 1] >> shift) | (_spread10(c[:, 2] >> shift) << np.uint64(1))
         | (_spread10(c[:, 3] >> shift) << np.uint64(2)))
    key = reflect(_global_keys(nbr)).astype(np.uint64)
    full = (key << np.uint64(37)) | ((c[:, 0].astype(np.uint64) & np.uint64(0x7F)) << np.uint64(30)) | m
    return full.view(np.int64)
