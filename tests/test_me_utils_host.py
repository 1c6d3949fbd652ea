"""Host-side ME.utils of the CUDA package (no GPU needed): batched_coordinates and sparse_collate with the reference's
own call pattern (data/alivev2.py:386-396 `collate_sparse`: lists of per-sample numpy / torch arrays, dtype=float32)
against the oracle package and a hand-written expectation; ragged, empty and integer inputs."""
import numpy as np
import pytest
import torch

import oracle.MinkowskiEngine as OME


def _samples(seed=3):
    rng = np.random.default_rng(seed)
    sizes = (7, 0, 4, 11)                       # ragged, one empty sample
    coords = [rng.normal(size=(n, 3)).astype(np.float32) * 50 for n in sizes]
    feats = [rng.random((n, 3)).astype(np.float32) for n in sizes]
    labels = [rng.integers(0, 3, size=(n,)).astype(np.int64) for n in sizes]
    return sizes, coords, feats, labels


def test_sparse_collate_reference_call_pattern(built_lib):
    import MinkowskiEngine as ME
    sizes, coords, feats, labels = _samples()
    # data/alivev2.py:391-396
    cb, fb, lb = ME.utils.sparse_collate(coords, feats, labels, dtype=torch.float32)
    ocb, ofb, olb = OME.utils.sparse_collate(coords, feats, labels, dtype=torch.float32)
    N = sum(sizes)
    assert cb.shape == (N, 4) and cb.dtype == torch.float32 and fb.shape == (N, 3) and lb.shape == (N,)
    assert torch.equal(cb, ocb) and torch.equal(fb, ofb) and torch.equal(lb, olb)
    # hand-written expectation: batch index column then the untouched coordinates, samples in order
    off = 0
    for b, n in enumerate(sizes):
        assert torch.all(cb[off:off + n, 0] == b)
        assert np.array_equal(cb[off:off + n, 1:].numpy(), coords[b])
        assert np.array_equal(fb[off:off + n].numpy(), feats[b])
        assert np.array_equal(lb[off:off + n].numpy(), labels[b])
        off += n
    # the offsets the reference derives from the labels (data/alivev2.py:405-407) index the collated rows
    assert off == N


def test_sparse_collate_int32_floors_and_no_labels(built_lib):
    import MinkowskiEngine as ME
    _, coords, feats, _ = _samples(5)
    cb, fb = ME.utils.sparse_collate(coords, feats)            # default dtype int32: floor, like ME
    ocb, ofb = OME.utils.sparse_collate(coords, feats)
    assert cb.dtype == torch.int32 and torch.equal(cb, ocb) and torch.equal(fb, ofb)
    want = np.concatenate([np.floor(c) for c in coords]).astype(np.int32)
    assert np.array_equal(cb[:, 1:].numpy(), want)
    # torch inputs and already-integer coordinates pass through unchanged
    ci = [torch.from_numpy(np.floor(c)).to(torch.int32) for c in coords]
    cb2, _ = ME.utils.sparse_collate(ci, [torch.from_numpy(f) for f in feats])
    assert torch.equal(cb2, cb)


def test_batched_coordinates_contract(built_lib):
    import MinkowskiEngine as ME
    _, coords, _, _ = _samples(7)
    for dt in (torch.int32, torch.float32):
        a = ME.utils.batched_coordinates(coords, dtype=dt)
        b = OME.utils.batched_coordinates(coords, dtype=dt)
        assert a.dtype == dt and torch.equal(a, b)
    with pytest.raises(ValueError):
        ME.utils.batched_coordinates(coords, dtype=torch.float64)
    assert ME.utils.batched_coordinates([]).shape == (0, 4)
