"""Oracle (TEST INFRASTRUCTURE) for the PointNet++ primitives: a dense PyTorch-CPU restatement of the semantics of
model/pointnet2_utils.py (farthest_point_sample :65-86, query_ball_point :89-110, the 3-NN weights of
PointNetFeaturePropagation.forward :283-292). Pinned by tests/golden/reference_pointnet2.npz, which holds outputs of
the reference's own functions (tests/golden/make_golden_pointnet2.py). Only tests/ may import this module."""
import torch


def pairwise_sqdist(a, b):
    """[B,N,3] x [B,M,3] -> [B,N,M] in the reference's expanded form -2<a,b> + |a|^2 + |b|^2 (:22-44)."""
    d = -2 * torch.matmul(a, b.transpose(1, 2))
    d = d + (a * a).sum(-1)[:, :, None]
    return d + (b * b).sum(-1)[:, None, :]


def fps(xyz, npoint, start):
    """xyz [B,N,3], start [B] -> [B,npoint] indices; running min distance to the chosen set, arg-max next."""
    B, N, _ = xyz.shape
    chosen = torch.zeros(B, npoint, dtype=torch.long)
    mind = torch.full((B, N), 1e10)
    cur = start.long().clone()
    rows = torch.arange(B)
    for s in range(npoint):
        chosen[:, s] = cur
        c = xyz[rows, cur][:, None, :]
        mind = torch.minimum(mind, ((xyz - c) ** 2).sum(-1))
        cur = mind.argmax(dim=1)
    return chosen


def ball_query(radius, nsample, xyz, new_xyz):
    """first nsample in-range indices in ascending order, padded with the first; N where a query has none."""
    B, N, _ = xyz.shape
    d = pairwise_sqdist(new_xyz, xyz)
    idx = torch.arange(N).expand(B, new_xyz.shape[1], N).clone()
    idx[d > radius * radius] = N
    idx = idx.sort(dim=-1)[0][:, :, :nsample]
    first = idx[:, :, :1].expand(-1, -1, nsample)
    return torch.where(idx == N, first, idx)


def three_nn(xyz1, xyz2):
    """-> (idx [B,N,3], weights [B,N,3], sqdist [B,N,3])"""
    d, idx = pairwise_sqdist(xyz1, xyz2).sort(dim=-1)
    d, idx = d[:, :, :3], idx[:, :, :3]
    r = 1.0 / (d + 1e-8)
    return idx, r / r.sum(-1, keepdim=True), d
