"""Mirror of model/pointnet2.py PointNet2SSG (:9-43), the reference's default key-point network
(config/default.yaml:179), on the GPU-native primitives of pointnet2_utils.

The network is described by two tables (set-abstraction levels going down, feature-propagation levels coming back up)
and built in the reference's construction order, so the module names - hence the state-dict keys - and, under the same
torch seed, the random initialisation are those of the reference class (tests/test_models_vs_reference.py)."""
import torch.nn as nn
import torch.nn.functional as F

from .pointnet2_utils import PointNetSetAbstraction, PointNetFeaturePropagation

# name, centroids, ball radius [m], samples per ball, feature channels coming in, MLP widths
_DOWN = (("sa1", 1024, 0.1, 32, None, (32, 32, 64)),
         ("sa2", 256, 0.2, 32, 64, (64, 64, 128)),
         ("sa3", 64, 0.4, 32, 128, (128, 128, 256)),
         ("sa4", 16, 0.8, 32, 256, (256, 256, 512)))
# name, channels coming in (skip + interpolated), MLP widths; fp4 joins levels 3 and 4, ... fp1 levels 0 and 1
_UP = (("fp4", 768, (256, 256)), ("fp3", 384, (256, 256)), ("fp2", 320, (256, 128)), ("fp1", 128, (128, 128, 128)))
_HEAD_WIDTH = 128


class PointNet2SSG(nn.Module):
    """single-scale-grouping PointNet++ segmentation net: input [B, C, N] (xyz first), output per-point logits
    [B, N, num_classes] and the coarsest level's features."""

    def __init__(self, num_classes=10, in_channels=3):
        super().__init__()
        for name, npoint, radius, nsample, feat_in, widths in _DOWN:
            cin = (in_channels if feat_in is None else feat_in) + 3   # grouped features are concatenated with xyz offsets
            self.add_module(name, PointNetSetAbstraction(npoint, radius, nsample, cin, list(widths), False))
        for name, cin, widths in _UP:
            self.add_module(name, PointNetFeaturePropagation(cin, list(widths)))
        self.conv1 = nn.Conv1d(_HEAD_WIDTH, _HEAD_WIDTH, 1)
        self.bn1 = nn.BatchNorm1d(_HEAD_WIDTH)
        self.drop1 = nn.Dropout(0.5)
        self.conv2 = nn.Conv1d(_HEAD_WIDTH, num_classes, 1)

    def forward(self, xyz):
        # level 0 = the input cloud (its features are the full input rows); levels 1..4 from the set abstractions
        coords, feats = [xyz[:, :3, :]], [xyz]
        for name, *_ in _DOWN:
            c, f = getattr(self, name)(coords[-1], feats[-1])
            coords.append(c)
            feats.append(f)
        coarsest = feats[-1]
        # propagate back up: level L receives the interpolation of level L + 1; level 0 has no skip features
        for level, (name, *_) in zip(range(len(_DOWN) - 1, -1, -1), _UP):
            skip = feats[level] if level > 0 else None
            feats[level] = getattr(self, name)(coords[level], coords[level + 1], skip, feats[level + 1])
        x = self.drop1(F.relu(self.bn1(self.conv1(feats[0]))))
        return self.conv2(x).permute(0, 2, 1), coarsest
