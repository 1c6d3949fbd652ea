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
    """model/pointnet2_utils.py:22-44 (kept for callers; the kernels below do not build this matrix)."""
    B, N, _ = src.shape
    _, M, _ = dst.shape
    dist = -2 * torch.matmul(src, dst.permute(0, 2, 1))
    dist += torch.sum(src ** 2, -1).view(B, N, 1)
    dist += torch.sum(dst ** 2, -1).view(B, 1, M)
    return dist


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


def sample_and_group(npoint, radius, nsample, xyz, points, returnfps=False):
    """model/pointnet2_utils.py:113-140."""
    B, N, C = xyz.shape
    S = npoint
    fps_idx = farthest_point_sample(xyz, npoint)
    new_xyz = index_points(xyz, fps_idx)
    idx = query_ball_point(radius, nsample, xyz, new_xyz)
    grouped_xyz = index_points(xyz, idx)
    grouped_xyz_norm = grouped_xyz - new_xyz.view(B, S, 1, C)
    if points is not None:
        new_points = torch.cat([grouped_xyz_norm, index_points(points, idx)], dim=-1)
    else:
        new_points = grouped_xyz_norm
    if returnfps:
        return new_xyz, new_points, grouped_xyz, fps_idx
    return new_xyz, new_points


def sample_and_group_all(xyz, points):
    """model/pointnet2_utils.py:143-161."""
    B, N, C = xyz.shape
    new_xyz = torch.zeros(B, 1, C, device=xyz.device)
    grouped_xyz = xyz.view(B, 1, N, C)
    if points is not None:
        new_points = torch.cat([grouped_xyz, points.view(B, 1, N, -1)], dim=-1)
    else:
        new_points = grouped_xyz
    return new_xyz, new_points


class PointNetSetAbstraction(nn.Module):
    """model/pointnet2_utils.py:164-204 (same attributes -> same state-dict keys)."""

    def __init__(self, npoint, radius, nsample, in_channel, mlp, group_all):
        super().__init__()
        self.npoint, self.radius, self.nsample = npoint, radius, nsample
        self.mlp_convs = nn.ModuleList()
        self.mlp_bns = nn.ModuleList()
        last_channel = in_channel
        for out_channel in mlp:
            self.mlp_convs.append(nn.Conv2d(last_channel, out_channel, 1))
            self.mlp_bns.append(nn.BatchNorm2d(out_channel))
            last_channel = out_channel
        self.group_all = group_all

    def forward(self, xyz, points):
        xyz = xyz.permute(0, 2, 1)
        if points is not None:
            points = points.permute(0, 2, 1)
        if self.group_all:
            new_xyz, new_points = sample_and_group_all(xyz, points)
        else:
            new_xyz, new_points = sample_and_group(self.npoint, self.radius, self.nsample, xyz, points)
        new_points = new_points.permute(0, 3, 2, 1)
        for conv, bn in zip(self.mlp_convs, self.mlp_bns):
            new_points = F.relu(bn(conv(new_points)))
        new_points = torch.max(new_points, 2)[0]
        return new_xyz.permute(0, 2, 1), new_points


class PointNetSetAbstractionMsg(nn.Module):
    """model/pointnet2_utils.py:207-262."""

    def __init__(self, npoint, radius_list, nsample_list, in_channel, mlp_list):
        super().__init__()
        self.npoint, self.radius_list, self.nsample_list = npoint, radius_list, nsample_list
        self.conv_blocks = nn.ModuleList()
        self.bn_blocks = nn.ModuleList()
        for mlp in mlp_list:
            convs, bns = nn.ModuleList(), nn.ModuleList()
            last_channel = in_channel + 3
            for out_channel in mlp:
                convs.append(nn.Conv2d(last_channel, out_channel, 1))
                bns.append(nn.BatchNorm2d(out_channel))
                last_channel = out_channel
            self.conv_blocks.append(convs)
            self.bn_blocks.append(bns)

    def forward(self, xyz, points):
        xyz = xyz.permute(0, 2, 1)
        if points is not None:
            points = points.permute(0, 2, 1)
        B, N, C = xyz.shape
        S = self.npoint
        new_xyz = index_points(xyz, farthest_point_sample(xyz, S))
        new_points_list = []
        for i, radius in enumerate(self.radius_list):
            group_idx = query_ball_point(radius, self.nsample_list[i], xyz, new_xyz)
            grouped_xyz = index_points(xyz, group_idx) - new_xyz.view(B, S, 1, C)
            if points is not None:
                grouped_points = torch.cat([index_points(points, group_idx), grouped_xyz], dim=-1)
            else:
                grouped_points = grouped_xyz
            grouped_points = grouped_points.permute(0, 3, 2, 1)
            for conv, bn in zip(self.conv_blocks[i], self.bn_blocks[i]):
                grouped_points = F.relu(bn(conv(grouped_points)))
            new_points_list.append(torch.max(grouped_points, 2)[0])
        return new_xyz.permute(0, 2, 1), torch.cat(new_points_list, dim=1)


class PointNetFeaturePropagation(nn.Module):
    """model/pointnet2_utils.py:265-318."""

    def __init__(self, in_channel, mlp):
        super().__init__()
        self.mlp_convs = nn.ModuleList()
        self.mlp_bns = nn.ModuleList()
        last_channel = in_channel
        for out_channel in mlp:
            self.mlp_convs.append(nn.Conv1d(last_channel, out_channel, 1))
            self.mlp_bns.append(nn.BatchNorm1d(out_channel))
            last_channel = out_channel

    def forward(self, xyz1, xyz2, points1, points2):
        xyz1 = xyz1.permute(0, 2, 1)
        xyz2 = xyz2.permute(0, 2, 1)
        points2 = points2.permute(0, 2, 1)
        B, N, C = xyz1.shape
        S = xyz2.shape[1]
        if S == 1:
            interpolated_points = points2.repeat(1, N, 1)
        else:
            idx, weight = three_nn(xyz1, xyz2)
            interpolated_points = torch.sum(index_points(points2, idx) * weight.view(B, N, 3, 1), dim=2)
        if points1 is not None:
            new_points = torch.cat([points1.permute(0, 2, 1), interpolated_points], dim=-1)
        else:
            new_points = interpolated_points
        new_points = new_points.permute(0, 2, 1)
        for conv, bn in zip(self.mlp_convs, self.mlp_bns):
            new_points = F.relu(bn(conv(new_points)))
        return new_points
