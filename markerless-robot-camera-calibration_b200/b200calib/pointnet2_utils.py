"""Drop-in for model/pointnet2_utils.py (same names, argument meaning and tensor layouts) with the O(N*M) dense
PyTorch primitives replaced by libb2me kernels:

  farthest_point_sample   npoint sequential torch steps              -> one CTA per cloud (b2me_fps)
  query_ball_point        [B,S,N] distance matrix + full sort        -> one warp per query (b2me_ball_query)
  3-NN interpolation      [B,N,S] distance matrix + full sort        -> one thread per point (b2me_three_nn)

The shared MLPs (1x1 Conv2d / Conv1d + BatchNorm + ReLU) and the gathers stay PyTorch library calls. To run the
reference's unchanged model/pointnet2.py on top of this module:  sys.modules["model.pointnet2_utils"] = this module.
CUDA only: there is no CPU fallback."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from MinkowskiEngine._lib import lib, check, ptr, stream, B2MEError
from MinkowskiEngine.core import _count


def _need_cuda(t):
    if not t.is_cuda:
        raise B2MEError("PointNet++ primitives run on a CUDA device only (no CPU fallback)")


def square_distance(src, dst):
    """model/pointnet2_utils.py:22-44: all pairwise squared distances [B,N,M] (kept for callers of the module; the
    kernels below never build this matrix). |s|^2 + |d|^2 - 2 s.d like the reference, so the rounding is the same."""
    cross = torch.bmm(src, dst.transpose(1, 2))
    return (-2 * cross + (src * src).sum(-1, keepdim=True)) + (dst * dst).sum(-1).unsqueeze(1)


def index_points(points, idx):
    """model/pointnet2_utils.py:47-62: points [B,N,C], idx [B,S] or [B,S,K] -> [B,S,(K,)C]."""
    B = points.shape[0]
    flat = idx.reshape(B, -1).long()
    out = torch.gather(points, 1, flat.unsqueeze(-1).expand(-1, -1, points.shape[-1]))
    return out.reshape(*idx.shape, points.shape[-1])


def farthest_point_sample(xyz, npoint, start=None):
    """model/pointnet2_utils.py:65-86. xyz [B,N,3] -> [B,npoint] int64. `start` [B]: first sample of every cloud;
    None draws it exactly like the reference (torch.randint on the default CPU generator)."""
    _need_cuda(xyz)
    B, N, _ = xyz.shape
    if start is None:
        start = torch.randint(0, N, (B,), dtype=torch.long)
    start = torch.as_tensor(start).to(xyz.device, torch.int32).contiguous()
    x = xyz.to(torch.float32).contiguous()
    out = torch.empty((B, npoint), dtype=torch.int32, device=xyz.device)
    check(lib.b2me_fps(ptr(x), B, N, npoint, ptr(start), ptr(out), stream()), "fps")
    _count(1)
    return out.long()


def query_ball_point(radius, nsample, xyz, new_xyz):
    """model/pointnet2_utils.py:89-110. -> group_idx [B,S,nsample] int64."""
    _need_cuda(xyz)
    B, N, _ = xyz.shape
    S = new_xyz.shape[1]
    out = torch.empty((B, S, nsample), dtype=torch.int32, device=xyz.device)
    check(lib.b2me_ball_query(ptr(xyz.to(torch.float32).contiguous()), ptr(new_xyz.to(torch.float32).contiguous()), B, N, S,
                              float(radius), nsample, ptr(out), stream()), "ball_query")
    _count(1)
    return out.long()


def three_nn(xyz1, xyz2):
    """the 3-NN step of PointNetFeaturePropagation.forward (:283-292): -> (idx [B,N,3] int64, weight [B,N,3] f32)."""
    _need_cuda(xyz1)
    B, N, _ = xyz1.shape
    S = xyz2.shape[1]
    idx = torch.empty((B, N, 3), dtype=torch.int32, device=xyz1.device)
    w = torch.empty((B, N, 3), dtype=torch.float32, device=xyz1.device)
    check(lib.b2me_three_nn(ptr(xyz1.to(torch.float32).contiguous()), ptr(xyz2.to(torch.float32).contiguous()), B, N, S,
                            ptr(idx), ptr(w), stream()), "three_nn")
    _count(1)
    return idx.long(), w


def _rows_last(t):
    """[B, C, N] (the layout the modules exchange) -> [B, N, C] (the layout the primitives take); None stays None."""
    return None if t is None else t.transpose(1, 2)


def _shared_mlp(conv_cls, bn_cls, cin, widths):
    """(convs, bns): 1x1 convolutions + batch norms of a shared MLP, as two ModuleLists (state-dict keys
    mlp_convs.i.* / mlp_bns.i.* of the reference modules); created conv, bn, conv, bn, ... like the reference does."""
    convs, bns = nn.ModuleList(), nn.ModuleList()
    for cout in widths:
        convs.append(conv_cls(cin, cout, 1))
        bns.append(bn_cls(cout))
        cin = cout
    return convs, bns


def _apply_mlp(x, convs, bns):
    for conv, bn in zip(convs, bns):
        x = F.relu(bn(conv(x)))
    return x


def sample_and_group(npoint, radius, nsample, xyz, points, returnfps=False):
    """model/pointnet2_utils.py:113-140: FPS centroids, a ball of `nsample` neighbours around each, neighbour
    coordinates relative to their centroid (+ the neighbours' features). xyz [B,N,3], points [B,N,D] or None ->
    new_xyz [B,npoint,3], new_points [B,npoint,nsample,3(+D)]."""
    centroid_idx = farthest_point_sample(xyz, npoint)
    new_xyz = index_points(xyz, centroid_idx)
    ball_idx = query_ball_point(radius, nsample, xyz, new_xyz)
    ball_xyz = index_points(xyz, ball_idx)
    new_points = ball_xyz - new_xyz.unsqueeze(2)
    if points is not None:
        new_points = torch.cat((new_points, index_points(points, ball_idx)), dim=-1)
    return (new_xyz, new_points, ball_xyz, centroid_idx) if returnfps else (new_xyz, new_points)


def sample_and_group_all(xyz, points):
    """model/pointnet2_utils.py:143-161: one group holding the whole cloud, centred on the origin."""
    B, N, C = xyz.shape
    everything = xyz.reshape(B, 1, N, C)
    if points is not None:
        everything = torch.cat((everything, points.reshape(B, 1, N, -1)), dim=-1)
    return xyz.new_zeros((B, 1, C)), everything


class PointNetSetAbstraction(nn.Module):
    """model/pointnet2_utils.py:164-204 (same attributes -> same state-dict keys): sample + group + shared MLP + max
    over each group. forward(xyz [B,3,N], points [B,D,N] | None) -> (new_xyz [B,3,S], new_points [B,mlp[-1],S])."""

    def __init__(self, npoint, radius, nsample, in_channel, mlp, group_all):
        super().__init__()
        self.npoint, self.radius, self.nsample, self.group_all = npoint, radius, nsample, group_all
        self.mlp_convs, self.mlp_bns = _shared_mlp(nn.Conv2d, nn.BatchNorm2d, in_channel, mlp)

    def forward(self, xyz, points):
        xyz, points = _rows_last(xyz), _rows_last(points)
        if self.group_all:
            new_xyz, grouped = sample_and_group_all(xyz, points)
        else:
            new_xyz, grouped = sample_and_group(self.npoint, self.radius, self.nsample, xyz, points)
        feats = _apply_mlp(grouped.permute(0, 3, 2, 1), self.mlp_convs, self.mlp_bns)   # [B, C, nsample, S]
        return new_xyz.transpose(1, 2), feats.max(dim=2).values


class PointNetSetAbstractionMsg(nn.Module):
    """model/pointnet2_utils.py:207-262: multi-scale grouping - one ball query + shared MLP per radius around the same
    FPS centroids, outputs concatenated along the channels."""

    def __init__(self, npoint, radius_list, nsample_list, in_channel, mlp_list):
        super().__init__()
        self.npoint, self.radius_list, self.nsample_list = npoint, radius_list, nsample_list
        self.conv_blocks, self.bn_blocks = nn.ModuleList(), nn.ModuleList()
        for widths in mlp_list:
            convs, bns = _shared_mlp(nn.Conv2d, nn.BatchNorm2d, in_channel + 3, widths)
            self.conv_blocks.append(convs)
            self.bn_blocks.append(bns)

    def forward(self, xyz, points):
        xyz, points = _rows_last(xyz), _rows_last(points)
        new_xyz = index_points(xyz, farthest_point_sample(xyz, self.npoint))
        scales = []
        for radius, nsample, convs, bns in zip(self.radius_list, self.nsample_list, self.conv_blocks, self.bn_blocks):
            ball_idx = query_ball_point(radius, nsample, xyz, new_xyz)
            grouped = index_points(xyz, ball_idx) - new_xyz.unsqueeze(2)
            if points is not None:
                grouped = torch.cat((index_points(points, ball_idx), grouped), dim=-1)   # features first, as the reference
            scales.append(_apply_mlp(grouped.permute(0, 3, 2, 1), convs, bns).max(dim=2).values)
        return new_xyz.transpose(1, 2), torch.cat(scales, dim=1)


class PointNetFeaturePropagation(nn.Module):
    """model/pointnet2_utils.py:265-318: features of the coarse level interpolated onto the fine level (inverse
    distance weights of the three nearest coarse points; a single coarse point is broadcast), concatenated with the
    fine level's own features, then a shared MLP. forward(xyz1 [B,3,N] fine, xyz2 [B,3,S] coarse, points1 [B,D1,N] |
    None, points2 [B,D2,S]) -> [B, mlp[-1], N]."""

    def __init__(self, in_channel, mlp):
        super().__init__()
        self.mlp_convs, self.mlp_bns = _shared_mlp(nn.Conv1d, nn.BatchNorm1d, in_channel, mlp)

    def forward(self, xyz1, xyz2, points1, points2):
        fine, coarse, coarse_feats = _rows_last(xyz1), _rows_last(xyz2), _rows_last(points2)
        n_fine = fine.shape[1]
        if coarse.shape[1] == 1:
            carried = coarse_feats.repeat(1, n_fine, 1)
        else:
            nn_idx, nn_w = three_nn(fine, coarse)
            carried = (index_points(coarse_feats, nn_idx) * nn_w.unsqueeze(-1)).sum(dim=2)
        if points1 is not None:
            carried = torch.cat((_rows_last(points1), carried), dim=-1)
        return _apply_mlp(carried.transpose(1, 2), self.mlp_convs, self.mlp_bns)
