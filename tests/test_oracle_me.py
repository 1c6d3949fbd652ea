"""Pin the CPU oracle of the MinkowskiEngine subset with independent dense PyTorch operators
(SURVEY.md §8c): conv3d / strided conv3d / conv_transpose3d on zero-filled grids, torch.unique."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import oracle.MinkowskiEngine as OME


def _random_cloud(n, extent, seed, batches=1):
    g = torch.Generator().manual_seed(seed)
    pts = [torch.rand(n, 3, generator=g) * extent - extent / 3 for _ in range(batches)]  # negative coords too
    return pts


def _dense(st, G, off):
    """scatter a SparseTensor onto a dense [B,C,G,G,G] grid (coords shifted by off, divided by its stride)."""
    C = st.C.long()
    ts = st.tensor_stride[0] or 1
    B = int(C[:, 0].max()) + 1
    d = torch.zeros((B, st.F.shape[1], G, G, G))
    idx = (C[:, 1:] + off) // ts
    d[C[:, 0], :, idx[:, 0], idx[:, 1], idx[:, 2]] = st.F
    return d, idx


def test_unique_first_occurrence_vs_torch_unique():
    g = torch.Generator().manual_seed(1)
    q = torch.randint(-5, 6, (4000, 4), generator=g)
    q[:, 0] = torch.randint(0, 3, (4000,), generator=g)
    uc, inv, first = OME.unique_first_occurrence(q.numpy())
    tu, tinv = torch.unique(q, dim=0, return_inverse=True)
    assert len(uc) == len(tu)
    assert np.array_equal(uc[inv], q.numpy())                 # inverse map is exact
    assert np.all(np.diff(first) > 0)                         # rows come in first-occurrence order
    for v in range(0, len(uc), 37):                           # `first` really is the first member
        assert first[v] == np.nonzero(inv == v)[0][0]


def test_voxel_mean_matches_float_mean():
    g = torch.Generator().manual_seed(2)
    pts = torch.rand(5000, 3, generator=g) * 10
    feats = torch.rand(5000, 3, generator=g) - 0.5
    tf = OME.TensorField(features=feats, coordinates=OME.utils.batched_coordinates([pts], dtype=torch.float32))
    st = tf.sparse()
    inv = tf.inverse_mapping
    V = st.F.shape[0]
    ref = torch.zeros(V, 3, dtype=torch.float64).index_add_(0, inv, feats.double())
    ref /= torch.bincount(inv, minlength=V).unsqueeze(1)
    assert torch.allclose(st.F.double(), ref, atol=1e-7)
    assert torch.equal(st.slice(tf).F, st.F[inv])


@pytest.mark.parametrize("seed", [3, 4])
def test_conv_k3_vs_dense_conv3d(seed):
    pts = _random_cloud(600, 14.0, seed, batches=2)
    feats = torch.randn(1200, 5, generator=torch.Generator().manual_seed(seed))
    st = OME.SparseTensor(feats, OME.utils.batched_coordinates(pts))
    conv = OME.MinkowskiConvolution(5, 7, kernel_size=3, dimension=3)
    out = conv(st)
    G, off = 24, 8
    d, idx = _dense(st, G, off)
    # ME kernel index: x fastest -> dense weight [Cout, Cin, kx, ky, kz] with k = kx + 3 ky + 9 kz
    W = conv.kernel.detach().view(3, 3, 3, 5, 7).permute(4, 3, 2, 1, 0)  # [kz,ky,kx,ci,co] -> [co,ci,kx,ky,kz]
    dd = F.conv3d(d, W, padding=1)
    C = out.C.long()
    got = dd[C[:, 0], :, idx[:, 0], idx[:, 1], idx[:, 2]]
    assert torch.allclose(out.F, got, atol=1e-4, rtol=1e-4)


def test_conv_k2s2_and_transpose_vs_dense():
    pts = _random_cloud(500, 12.0, 5, batches=1)
    feats = torch.randn(500, 4, generator=torch.Generator().manual_seed(5))
    st = OME.SparseTensor(feats, OME.utils.batched_coordinates(pts))
    down = OME.MinkowskiConvolution(4, 6, kernel_size=2, stride=2, dimension=3)
    up = OME.MinkowskiConvolutionTranspose(6, 3, kernel_size=2, stride=2, dimension=3)
    y = down(st)
    assert y.tensor_stride == [2, 2, 2]
    G, off = 24, 8   # off even so that floor(c/2)*2 aligns with dense stride-2 windows
    d, _ = _dense(st, G, off)
    Wd = down.kernel.detach().view(2, 2, 2, 4, 6).permute(4, 3, 2, 1, 0)
    dd = F.conv3d(d, Wd, stride=2)
    Cy = y.C.long()
    iy = (Cy[:, 1:] + off) // 2
    assert torch.allclose(y.F, dd[Cy[:, 0], :, iy[:, 0], iy[:, 1], iy[:, 2]], atol=1e-4, rtol=1e-4)
    # every coarse coordinate is the floor-aligned parent of some fine voxel, each exactly once
    parents = torch.unique(torch.div(st.C[:, 1:], 2, rounding_mode="floor") * 2, dim=0)
    assert len(parents) == len(Cy)
    # transposed conv back onto the cached fine map
    z = up(y)
    assert z.tensor_stride == [1, 1, 1] and z.F.shape[0] == st.F.shape[0]
    dy = torch.zeros((1, 6, G // 2, G // 2, G // 2))
    dy[Cy[:, 0], :, iy[:, 0], iy[:, 1], iy[:, 2]] = y.F
    Wu = up.kernel.detach().view(2, 2, 2, 6, 3).permute(3, 4, 2, 1, 0)   # conv_transpose3d: [Cin, Cout, kx,ky,kz]
    dz = F.conv_transpose3d(dy, Wu, stride=2)
    Cz = z.C.long()
    iz = Cz[:, 1:] + off
    assert torch.allclose(z.F, dz[Cz[:, 0], :, iz[:, 0], iz[:, 1], iz[:, 2]], atol=1e-4, rtol=1e-4)


def test_unet_levels_consistent():
    """four stride-2 levels: tensor strides 1..16 and the decoder lands on the encoder's maps (cat succeeds)."""
    from b200calib.models import make_models
    M = make_models(OME)
    torch.manual_seed(0)
    net = M.MinkUNet(3, 8, variant="MinkUNet14A").eval()
    pts = torch.rand(3000, 3) * torch.tensor([40.0, 40.0, 3.0])
    tf = OME.TensorField(features=torch.rand(3000, 3), coordinates=OME.utils.batched_coordinates([pts], dtype=torch.float32))
    x = tf.sparse()
    with torch.no_grad():
        y = net(x)
    assert y.F.shape == (x.F.shape[0], 8)
    keys = sorted(k._ts for k in x.coordinate_manager.levels)
    assert keys == [1, 2, 4, 8, 16]


def test_sparse_quantize_labels_and_division():
    rng = np.random.default_rng(0)
    pts = rng.random((2000, 3)).astype(np.float32)
    feats = rng.random((2000, 3)).astype(np.float32)
    labels = rng.integers(0, 3, 2000).astype(np.int32)
    c, f, l = OME.utils.sparse_quantize(pts, feats, labels, quantization_size=0.1, ignore_label=-100)
    assert c.dtype == np.int32 and len(c) == len(f) == len(l)
    q = np.floor(pts / 0.1).astype(np.int32)
    assert len(np.unique(q, axis=0)) == len(c)
    for v in range(0, len(c), 11):
        members = np.all(q == c[v], axis=1)
        ls = np.unique(labels[members])
        assert l[v] == (ls[0] if len(ls) == 1 else -100)
        assert np.array_equal(f[v], feats[np.nonzero(members)[0][0]])


def test_batched_coordinates():
    a, b = torch.rand(5, 3) * 9 - 4, torch.rand(2, 3)
    bc = OME.utils.batched_coordinates([a, b], dtype=torch.float32)
    assert bc.shape == (7, 4) and torch.equal(bc[:, 0], torch.tensor([0, 0, 0, 0, 0, 1, 1.0]))
    bi = OME.utils.batched_coordinates([a, b])
    assert bi.dtype == torch.int32 and torch.equal(bi[:5, 1:], torch.floor(a).int())
