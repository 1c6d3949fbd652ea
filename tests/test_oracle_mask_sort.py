"""oracle/mask_sort.py (the NumPy statement of the K3b row-order keys; the device kernels are compared with it in
tests/test_gpu_configs.py): properties of the definitions themselves."""
import numpy as np

from oracle import mask_sort as MS


def _map(rng, V, K=27):
    nbr = rng.integers(0, V, (V, K)).astype(np.int32)
    nbr[rng.random((V, K)) < np.linspace(0.25, 0.97, K)] = -1
    nbr[:, K // 2] = np.arange(V)
    return nbr


def test_keys_group_equal_masks_and_cut_passes():
    rng = np.random.default_rng(7)
    nbr = _map(rng, 6000)
    masks = ((nbr >= 0) << np.arange(27)).sum(1)
    coords = np.concatenate((rng.integers(0, 3, (6000, 1)), rng.integers(-300, 700, (6000, 3))), 1).astype(np.int32)
    p_nat = MS.passes(nbr, np.arange(len(nbr)))
    for keys in (MS.keys_one_level(nbr), MS.keys_two_level(nbr), MS.keys_morton(nbr, coords, 2)):
        order = np.argsort(keys, kind="stable")
        assert np.array_equal(np.sort(order), np.arange(len(nbr)))
        # rows with the same neighbour pattern are contiguous in the sorted order
        m = masks[order]
        change = np.flatnonzero(m[1:] != m[:-1]) + 1
        assert len(change) + 1 == len(np.unique(masks))
        assert MS.passes(nbr, order) < 0.85 * p_nat
    # the rarest offset is the most significant bit of the one-level key (before reflection it splits the order in two)
    counts = (nbr >= 0).sum(0)
    rare = int(np.argmin(counts))
    order = np.argsort(MS.keys_one_level(nbr), kind="stable")
    has = (nbr[order, rare] >= 0)
    assert np.all(has[np.argmax(has):]) or np.all(~has[np.argmin(has):]) or has.sum() == 0


def test_reflect_is_a_gray_code():
    x = np.arange(1 << 12, dtype=np.uint32)
    g = MS.reflect(x)
    assert len(np.unique(g)) == len(x)                       # a permutation of the codes
    # consecutive SORTED reflected keys differ in exactly one mask bit
    inv = np.argsort(g)
    d = x[inv][1:] ^ x[inv][:-1]
    assert np.all((d & (d - 1)) == 0)
