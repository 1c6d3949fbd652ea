"""Host-side mirror of the reference's sparse networks, for machines where the reference tree is not
present (the GPU box). Same topology, attribute names and state-dict keys as

    model/backbone/minkunet.py:52-187   MinkUNetBase (+ variants :189-251)
    model/backbone/resnet.py:86-127     weight_initialization / _make_layer
    model/robotnet_segmentation.py:35-64, model/robotnet_vote.py:36-71   trunk + 256->1024->C head
    model/robotnet_encode.py:37-117     encoder-only trunk + BN/ReLU + global average pool + MLP

so that a checkpoint of the reference loads by key, and tests/test_models_vs_reference.py checks (in the
authoring container, where /root/reference exists) that the unchanged reference files produce identical
state-dict keys/shapes and identical outputs on the same ME implementation.

The classes are built against an ME implementation passed in (`make_models(ME)`): the CUDA package in
production, the CPU oracle in tests.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

# variant name -> (block kind, layers per stage, planes per stage)
UNET_VARIANTS = {
    "MinkUNet14": ("basic", (1,) * 8, (32, 64, 128, 256, 256, 128, 96, 96)),
    "MinkUNet14A": ("basic", (1,) * 8, (32, 64, 128, 256, 128, 128, 96, 96)),
    "MinkUNet14B": ("basic", (1,) * 8, (32, 64, 128, 256, 128, 128, 128, 128)),
    "MinkUNet14C": ("basic", (1,) * 8, (32, 64, 128, 256, 192, 192, 128, 128)),
    "MinkUNet14D": ("basic", (1,) * 8, (32, 64, 128, 256, 384, 384, 384, 384)),
    "MinkUNet18": ("basic", (2,) * 8, (32, 64, 128, 256, 256, 128, 96, 96)),
    "MinkUNet18A": ("basic", (2,) * 8, (32, 64, 128, 256, 128, 128, 96, 96)),
    "MinkUNet18B": ("basic", (2,) * 8, (32, 64, 128, 256, 128, 128, 128, 128)),
    "MinkUNet18D": ("basic", (2,) * 8, (32, 64, 128, 256, 384, 384, 384, 384)),
    "MinkUNet34": ("basic", (2, 3, 4, 6, 2, 2, 2, 2), (32, 64, 128, 256, 256, 128, 96, 96)),
    "MinkUNet34A": ("basic", (2, 3, 4, 6, 2, 2, 2, 2), (32, 64, 128, 256, 256, 128, 64, 64)),
    "MinkUNet34B": ("basic", (2, 3, 4, 6, 2, 2, 2, 2), (32, 64, 128, 256, 256, 128, 64, 32)),
    "MinkUNet34C": ("basic", (2, 3, 4, 6, 2, 2, 2, 2), (32, 64, 128, 256, 256, 128, 96, 96)),
    "MinkUNet50": ("bottleneck", (2, 3, 4, 6, 2, 2, 2, 2), (32, 64, 128, 256, 256, 128, 96, 96)),
    "MinkUNet101": ("bottleneck", (2, 3, 4, 23, 2, 2, 2, 2), (32, 64, 128, 256, 256, 128, 96, 96)),
}

# encoder stage i: stride-2 conv name, its BN, residual stage name
_ENC = [("conv1p1s2", "bn1", "block1"), ("conv2p2s2", "bn2", "block2"),
        ("conv3p4s2", "bn3", "block3"), ("conv4p8s2", "bn4", "block4")]
_DEC = [("convtr4p16s2", "bntr4", "block5"), ("convtr5p8s2", "bntr5", "block6"),
        ("convtr6p4s2", "bntr6", "block7"), ("convtr7p2s2", "bntr7", "block8")]
INIT_DIM = 32


def make_models(ME):
    """-> namespace of classes bound to the given MinkowskiEngine implementation."""
    blocks = {"basic": ME.modules.resnet_block.BasicBlock, "bottleneck": ME.modules.resnet_block.Bottleneck}

    class MinkUNet(nn.Module):
        def __init__(self, in_channels, out_channels, D=3, variant="MinkUNet18D"):
            super().__init__()
            kind, layers, planes = UNET_VARIANTS[variant]
            self.BLOCK = blocks[kind]
            self.LAYERS, self.PLANES, self.D = layers, planes, D
            exp = self.BLOCK.expansion
            self.inplanes = INIT_DIM
            self.conv0p1s1 = ME.MinkowskiConvolution(in_channels, self.inplanes, kernel_size=3, dimension=D)
            self.bn0 = ME.MinkowskiBatchNorm(self.inplanes)
            for i, (cname, bname, blk) in enumerate(_ENC):
                setattr(self, cname, ME.MinkowskiConvolution(self.inplanes, self.inplanes, kernel_size=2, stride=2,
                                                             dimension=D))
                setattr(self, bname, ME.MinkowskiBatchNorm(self.inplanes))
                setattr(self, blk, self._stage(planes[i], layers[i]))
            skip_channels = [planes[2] * exp, planes[1] * exp, planes[0] * exp, INIT_DIM]
            for i, (cname, bname, blk) in enumerate(_DEC):
                p = planes[4 + i]
                setattr(self, cname, ME.MinkowskiConvolutionTranspose(self.inplanes, p, kernel_size=2, stride=2,
                                                                      dimension=D))
                setattr(self, bname, ME.MinkowskiBatchNorm(p))
                self.inplanes = p + skip_channels[i]
                setattr(self, blk, self._stage(p, layers[4 + i]))
            self.final = ME.MinkowskiConvolution(planes[7] * exp, out_channels, kernel_size=1, bias=True, dimension=D)
            self.relu = ME.MinkowskiReLU(inplace=True)
            self._init_weights()

        def _stage(self, planes, n):
            exp = self.BLOCK.expansion
            down = None
            if self.inplanes != planes * exp:
                down = nn.Sequential(
                    ME.MinkowskiConvolution(self.inplanes, planes * exp, kernel_size=1, stride=1, dimension=self.D),
                    ME.MinkowskiBatchNorm(planes * exp))
            mods = [self.BLOCK(self.inplanes, planes, stride=1, dilation=1, downsample=down, dimension=self.D)]
            self.inplanes = planes * exp
            mods += [self.BLOCK(self.inplanes, planes, stride=1, dilation=1, dimension=self.D) for _ in range(1, n)]
            return nn.Sequential(*mods)

        def _init_weights(self):
            for m in self.modules():
                if isinstance(m, ME.MinkowskiConvolution):  # transposed convs keep the default uniform init
                    ME.utils.kaiming_normal_(m.kernel, mode="fan_out", nonlinearity="relu")
                if isinstance(m, ME.MinkowskiBatchNorm):
                    nn.init.constant_(m.bn.weight, 1)
                    nn.init.constant_(m.bn.bias, 0)

        def encode(self, x):
            """stem + four stride-2 stages; returns (deepest features, skips fine->coarse)."""
            out = self.relu(self.bn0(self.conv0p1s1(x)))
            skips = [out]
            for i, (cname, bname, blk) in enumerate(_ENC):
                out = self.relu(getattr(self, bname)(getattr(self, cname)(out)))
                out = getattr(self, blk)(out)
                if i < 3:
                    skips.append(out)
            return out, skips

        def forward_except_final(self, x):
            out, skips = self.encode(x)
            for (cname, bname, blk), skip in zip(_DEC, reversed(skips)):
                out = self.relu(getattr(self, bname)(getattr(self, cname)(out)))
                out = getattr(self, blk)(ME.cat(out, skip))
            return out

        def forward(self, x):
            return self.final(self.forward_except_final(x))

    class _RobotNetPerPoint(MinkUNet):
        """MinkUNet -> LeakyReLU -> Linear 256->1024 -> LeakyReLU -> Linear 1024->num_classes."""
        name = "robotnet"

        def __init__(self, in_channels, out_channels=256, D=3, num_classes=3, variant="MinkUNet18D"):
            super().__init__(in_channels, out_channels, D, variant=variant)
            self.leaky_relu = ME.MinkowskiLeakyReLU()
            self.regression = nn.Sequential(ME.MinkowskiOps.MinkowskiLinear(256, 1024), ME.MinkowskiLeakyReLU(),
                                            ME.MinkowskiOps.MinkowskiLinear(1024, num_classes))
            self.sigm = ME.MinkowskiSigmoid()

        def forward(self, x):
            if isinstance(x, tuple):
                x = x[0]
            return self.regression(self.leaky_relu(MinkUNet.forward(self, x)))

    class RobotNetSegmentation(_RobotNetPerPoint):
        pass

    class RobotNetVote(_RobotNetPerPoint):
        def __init__(self, in_channels, out_channels=256, D=3, num_classes=2, variant="MinkUNet18D"):
            super().__init__(in_channels, out_channels, D, num_classes=num_classes, variant=variant)

    class RobotNetEncode(MinkUNet):
        """encoder-only trunk -> BN+ReLU -> global average pool -> Linear 2048 -> LeakyReLU -> Linear out."""
        name = "robotnet"

        def __init__(self, in_channels, out_channels, D=3, variant="MinkUNet18D", use_joint_angles=False,
                     quantization_size=0.01, voxelize_position=False):
            super().__init__(in_channels, out_channels, D, variant=variant)
            self.global_pool = ME.MinkowskiGlobalAvgPooling()
            self.leaky_relu = nn.LeakyReLU()
            self.final_bn = ME.MinkowskiBatchNorm(out_channels)
            width = self.PLANES[3] * self.BLOCK.expansion
            self.output_layer = nn.Sequential(ME.MinkowskiBatchNorm(width), self.relu)
            self.use_joint_angles = use_joint_angles
            self.pose_regression_input_size = width + (9 if use_joint_angles else 0)
            self.pose_regression = nn.Sequential(nn.Linear(self.pose_regression_input_size, 2048), nn.LeakyReLU(),
                                                 nn.Linear(2048, out_channels))
            self.quantization_size = quantization_size
            self.voxelize_position = voxelize_position

        def forward(self, x):  # quaternion columns are WXYZ
            joint_angles = None
            if isinstance(x, tuple):
                x, joint_angles = x
            deep, _ = self.encode(x)
            pooled = self.global_pool(self.output_layer(deep)).features.float()
            if self.use_joint_angles:
                pooled = torch.cat((pooled, joint_angles), dim=1)
            out = self.pose_regression(pooled)
            out[:, 7:] = torch.sigmoid(out[:, 7:])
            if not self.training:
                out[:, 3:7] = F.normalize(out[:, 3:7], p=2, dim=1)
                if self.voxelize_position:
                    out[:, :3] *= self.quantization_size
            return out

    class RobotNet(MinkUNet):
        """full UNet -> BN+ReLU -> global MAX pool -> Linear 2048 -> LeakyReLU -> Linear out
        (model/robotnet.py:37-83), the non-encode_only rotation regressor."""
        name = "robotnet"

        def __init__(self, in_channels, out_channels, D=3, variant="MinkUNet18D", use_joint_angles=False):
            super().__init__(in_channels, out_channels, D, variant=variant)
            self.global_pool = ME.MinkowskiGlobalMaxPooling()
            self.leaky_relu = ME.MinkowskiLeakyReLU(inplace=False)
            self.final_bn = ME.MinkowskiBatchNorm(out_channels)
            width = self.PLANES[-1] * self.BLOCK.expansion
            self.output_layer = nn.Sequential(ME.MinkowskiBatchNorm(width), self.relu)
            self.use_joint_angles = use_joint_angles
            self.pose_regression_input_size = width + (9 if use_joint_angles else 0)
            self.pose_regression = nn.Sequential(nn.Linear(self.pose_regression_input_size, 2048), nn.LeakyReLU(),
                                                 nn.Linear(2048, out_channels))

        def forward(self, x):  # quaternion columns are WXYZ
            joint_angles = None
            if isinstance(x, tuple):
                x, joint_angles = x
            feats = self.output_layer(self.forward_except_final(x))
            pooled = self.global_pool(feats).features.float()
            if self.use_joint_angles:
                pooled = torch.cat((pooled, joint_angles), dim=1)
            out = self.pose_regression(pooled)
            out[:, 7:] = torch.sigmoid(out[:, 7:])
            if not self.training:
                out[:, 3:7] = F.normalize(out[:, 3:7], p=2, dim=1)
            return out

    class AliveUNet(MinkUNet):
        """model/backbone/aliveunet.py:45-275: the seven-level U-Net the reference falls back to when
        STRUCTURE.backbone names no MinkUNet variant. PLANES = m x (1..7, 7..1) (m = STRUCTURE.m, 32 in
        config/default.yaml), `block_reps` blocks per stage, BasicBlock or Bottleneck; stride-2 convolutions down to
        tensor stride 128, transposed convolutions back up, every decoder stage takes the concatenation with the
        encoder output of its stride. Same attribute names / state-dict keys as the reference class. Its forward
        returns the last decoder stage (`final` is constructed but not applied, aliveunet.py:264-265)."""
        ENC = [("conv1p1s2", "bn1", "block1"), ("conv2p2s2", "bn2", "block2"), ("conv3p4s2", "bn3", "block3"),
               ("conv4p8s2", "bn4", "block4"), ("conv5p16s2", "bn5", "block5"), ("conv6p32s2", "bn6", "block6"),
               ("conv7p64s2", "bn7", "block7")]
        DEC = [("convtr7", "bntr7", "block8"), ("convtr8", "bntr8", "block9"), ("convtr9", "bntr9", "block10"),
               ("convtr10", "bntr10", "block11"), ("convtr11", "bntr11", "block12"), ("convtr12", "bntr12", "block13"),
               ("convtr13", "bntr13", "block14")]

        def __init__(self, in_channels, out_channels, D=3, m=32, block_reps=1, bottleneck=False):
            nn.Module.__init__(self)
            self.BLOCK = blocks["bottleneck" if bottleneck else "basic"]
            planes = tuple(i * m for i in (list(range(1, 8)) + list(range(7, 0, -1))))
            self.PLANES, self.LAYERS, self.D = planes, (block_reps,) * len(planes), D
            exp = self.BLOCK.expansion
            self.inplanes = INIT_DIM
            self.conv0p1s1 = ME.MinkowskiConvolution(in_channels, self.inplanes, kernel_size=3, dimension=D)
            self.bn0 = ME.MinkowskiBatchNorm(self.inplanes)
            for i, (cname, bname, blk) in enumerate(self.ENC):
                setattr(self, cname, ME.MinkowskiConvolution(self.inplanes, self.inplanes, kernel_size=2, stride=2,
                                                             dimension=D))
                setattr(self, bname, ME.MinkowskiBatchNorm(self.inplanes))
                setattr(self, blk, self._stage(planes[i], block_reps))
            # decoder stage j (block 8 + j): transposed conv to planes[7 + j] channels (the reference passes
            # PLANES[7] for the first and PLANES[8 + j - 1] after it, aliveunet.py:118-168), then the stage on
            # planes[8 + j] + skip channels; the last stage repeats planes[13] on the stem output
            tr_out = [planes[7]] + [planes[8 + j] for j in range(6)]
            stage_p = [planes[8 + j] for j in range(6)] + [planes[13]]
            # the reference sizes stage j for planes[8 + j] + planes[6 - j] * expansion input channels
            # (aliveunet.py:123,131,...): equal to what forward concatenates (planes[7 + j] + planes[5 - j] * expansion)
            # only for BasicBlock - with Bottleneck the reference class constructs but cannot run; mirrored as is
            skip_ch = [planes[6 - j] * exp for j in range(6)] + [INIT_DIM]
            for j, (cname, bname, blk) in enumerate(self.DEC):
                setattr(self, cname, ME.MinkowskiConvolutionTranspose(self.inplanes, tr_out[j], kernel_size=2, stride=2,
                                                                      dimension=D))
                setattr(self, bname, ME.MinkowskiBatchNorm(tr_out[j]))
                self.inplanes = stage_p[j] + skip_ch[j]
                setattr(self, blk, self._stage(stage_p[j], block_reps))
            self.final = ME.MinkowskiConvolution(planes[13] * exp, out_channels, kernel_size=1, bias=True, dimension=D)
            self.relu = ME.MinkowskiReLU(inplace=True)
            self._init_weights()

        def forward(self, x):
            out = self.relu(self.bn0(self.conv0p1s1(x)))
            skips = [out]
            for i, (cname, bname, blk) in enumerate(self.ENC):
                out = self.relu(getattr(self, bname)(getattr(self, cname)(out)))
                out = getattr(self, blk)(out)
                if i < 6:
                    skips.append(out)
            for (cname, bname, blk), skip in zip(self.DEC, reversed(skips)):
                out = self.relu(getattr(self, bname)(getattr(self, cname)(out)))
                out = getattr(self, blk)(ME.cat(out, skip))
            return out

    ns = type("Models", (), {})()
    ns.AliveUNet = AliveUNet
    ns.MinkUNet = MinkUNet
    ns.RobotNetSegmentation = RobotNetSegmentation
    ns.RobotNetVote = RobotNetVote
    ns.RobotNetEncode = RobotNetEncode
    ns.RobotNet = RobotNet
    return ns


def randomize_bn_stats(model, seed=13):
    """non-trivial eval-mode BatchNorm statistics for random-init tests:
    running_mean ~ N(0, 0.1), running_var ~ U[0.5, 1.5], gamma ~ U[0.8, 1.2], beta ~ N(0, 0.05)."""
    g = torch.Generator().manual_seed(seed)
    for m in model.modules():
        if isinstance(m, nn.BatchNorm1d):
            with torch.no_grad():
                m.running_mean.copy_(torch.randn(m.num_features, generator=g) * 0.1)
                m.running_var.copy_(torch.rand(m.num_features, generator=g) + 0.5)
                m.weight.copy_(torch.rand(m.num_features, generator=g) * 0.4 + 0.8)
                m.bias.copy_(torch.randn(m.num_features, generator=g) * 0.05)
    return model
