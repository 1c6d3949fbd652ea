"""Oracle residual blocks (ME modules/resnet_block.py semantics)."""
import torch.nn as nn

from .. import MinkowskiConvolution as Conv, MinkowskiBatchNorm as BN, MinkowskiReLU as ReLU


class _Block(nn.Module):
    def _tail(self, x, out):
        res = self.downsample(x) if self.downsample is not None else x
        out = out + res
        return self.relu(out)


class BasicBlock(_Block):
    expansion = 1

    def __init__(self, inplanes, planes, stride=1, dilation=1, downsample=None, bn_momentum=0.1, dimension=-1):
        super().__init__()
        self.conv1 = Conv(inplanes, planes, kernel_size=3, stride=stride, dilation=dilation, dimension=dimension)
        self.norm1 = BN(planes, momentum=bn_momentum)
        self.conv2 = Conv(planes, planes, kernel_size=3, stride=1, dilation=dilation, dimension=dimension)
        self.norm2 = BN(planes, momentum=bn_momentum)
        self.relu = ReLU(inplace=True)
        self.downsample = downsample

    def forward(self, x):
        out = self.relu(self.norm1(self.conv1(x)))
        out = self.norm2(self.conv2(out))
        return self._tail(x, out)


class Bottleneck(_Block):
    expansion = 4

    def __init__(self, inplanes, planes, stride=1, dilation=1, downsample=None, bn_momentum=0.1, dimension=-1):
        super().__init__()
        self.conv1 = Conv(inplanes, planes, kernel_size=1, dimension=dimension)
        self.norm1 = BN(planes, momentum=bn_momentum)
        self.conv2 = Conv(planes, planes, kernel_size=3, stride=stride, dilation=dilation, dimension=dimension)
        self.norm2 = BN(planes, momentum=bn_momentum)
        self.conv3 = Conv(planes, planes * 4, kernel_size=1, dimension=dimension)
        self.norm3 = BN(planes * 4, momentum=bn_momentum)
        self.relu = ReLU(inplace=True)
        self.downsample = downsample

    def forward(self, x):
        out = self.relu(self.norm1(self.conv1(x)))
        out = self.relu(self.norm2(self.conv2(out)))
        out = self.norm3(self.conv3(out))
        return self._tail(x, out)
