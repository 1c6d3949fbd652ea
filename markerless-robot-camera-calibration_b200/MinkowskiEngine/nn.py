"""nn.Module layer of the B200 MinkowskiEngine-compatible package: same class names, constructor
arguments, parameter names/shapes and state-dict keys as MinkowskiEngine 0.5.4 for the subset the reference
uses, so model/backbone/minkunet.py, model/backbone/resnet.py and model/robotnet_*.py run
unchanged."""
import math

import torch
import torch.nn as nn

from . import _lib, ops
from .core import SparseTensor, TensorField


def _scalar(v, name):
    if isinstance(v, (list, tuple)):
        if len(set(v)) != 1:
            raise NotImplementedError(f"anisotropic {name}={v}")
        v = v[0]
    return int(v)


class MinkowskiModuleBase(nn.Module):
    pass


class _ConvBase(MinkowskiModuleBase):
    is_transpose = False

    def __init__(self, in_channels, out_channels, kernel_size=-1, stride=1, dilation=1, bias=False,
                 kernel_generator=None, expand_coordinates=False, convolution_mode=None, dimension=None):
        super().__init__()
        if dimension is None:
            raise ValueError("dimension must be given")
        if dimension != 3:
            raise NotImplementedError("only 3-D sparse convolution is implemented")
        if kernel_generator is not None or expand_coordinates:
            raise NotImplementedError("kernel_generator / expand_coordinates are outside the hot-path subset")
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.kernel_size = _scalar(kernel_size, "kernel_size")
        self.stride = _scalar(stride, "stride")
        self.dilation = _scalar(dilation, "dilation")
        self.dimension = dimension
        self.kernel_volume = self.kernel_size ** dimension
        shape = (self.kernel_volume, in_channels, out_channels)
        if self.kernel_volume == 1:
            shape = (in_channels, out_channels)
        self.kernel = nn.Parameter(torch.empty(*shape, dtype=torch.float32))
        self.bias = nn.Parameter(torch.empty(1, out_channels, dtype=torch.float32)) if bias else None
        self.reset_parameters()

    def reset_parameters(self):
        with torch.no_grad():
            n = (self.out_channels if self.is_transpose else self.in_channels) * self.kernel_volume
            stdv = 1.0 / math.sqrt(n)
            self.kernel.uniform_(-stdv, stdv)
            if self.bias is not None:
                self.bias.uniform_(-stdv, stdv)

    def forward(self, input, coordinates=None):
        if coordinates is not None:
            raise NotImplementedError("explicit output coordinates")
        if not isinstance(input, SparseTensor):
            raise TypeError("MinkowskiConvolution expects a SparseTensor")
        if input.num_channels != self.in_channels:
            raise ValueError(f"expected {self.in_channels} input channels, got {input.num_channels}")
        return ops.conv_forward(self, input)

    def extra_repr(self):
        return (f"in={self.in_channels}, out={self.out_channels}, kernel_size={self.kernel_size}, "
                f"stride={self.stride}, dilation={self.dilation}")


class MinkowskiConvolution(_ConvBase):
    is_transpose = False


class MinkowskiConvolutionTranspose(_ConvBase):
    is_transpose = True


class MinkowskiBatchNorm(MinkowskiModuleBase):
    def __init__(self, num_features, eps=1e-5, momentum=0.1, affine=True, track_running_stats=True):
        super().__init__()
        self.bn = nn.BatchNorm1d(num_features, eps=eps, momentum=momentum, affine=affine,
                                 track_running_stats=track_running_stats)

    def forward(self, input):
        if self.training or not self.bn.track_running_stats:
            raise NotImplementedError("MinkowskiBatchNorm: only eval mode with running statistics is implemented "
                                      "(inference hot path)")
        scale, shift = ops._bn_affine(self.bn)
        return ops.affine_forward(input, scale, shift)


class MinkowskiReLU(MinkowskiModuleBase):
    def __init__(self, inplace=False):
        super().__init__()

    def forward(self, input):
        return ops.act_forward(input, _lib.ACT_RELU)


class MinkowskiLeakyReLU(MinkowskiModuleBase):
    def __init__(self, negative_slope=0.01, inplace=False):
        super().__init__()
        self.negative_slope = negative_slope

    def forward(self, input):
        return ops.act_forward(input, _lib.ACT_LEAKY, float(self.negative_slope))


class MinkowskiSigmoid(MinkowskiModuleBase):
    def forward(self, input):
        return input._child(features=torch.sigmoid(input.F.float()))


class MinkowskiLinear(MinkowskiModuleBase):
    def __init__(self, in_features, out_features, bias=True):
        super().__init__()
        self.linear = nn.Linear(in_features, out_features, bias=bias)

    def forward(self, input):
        if isinstance(input, TensorField):
            raise NotImplementedError("MinkowskiLinear on a TensorField")
        return ops.linear_forward(self, input)


class MinkowskiGlobalAvgPooling(MinkowskiModuleBase):
    def __init__(self, mode=None):
        super().__init__()

    def forward(self, input):
        return ops.global_pool(input, 0)


class MinkowskiGlobalMaxPooling(MinkowskiModuleBase):
    def __init__(self, mode=None):
        super().__init__()

    def forward(self, input):
        return ops.global_pool(input, 1)


MinkowskiGlobalPooling = MinkowskiGlobalAvgPooling


def _placeholder(name):
    class _P(MinkowskiModuleBase):
        def __init__(self, *a, **k):
            super().__init__()

        def forward(self, *a, **k):
            raise NotImplementedError(f"{name} is outside the inference hot-path subset (SURVEY.md §8b)")

    _P.__name__ = name
    _P.__qualname__ = name
    return _P


# referenced only inside never-instantiated classes of model/backbone/resnet.py
MinkowskiInstanceNorm = _placeholder("MinkowskiInstanceNorm")
MinkowskiMaxPooling = _placeholder("MinkowskiMaxPooling")
MinkowskiAvgPooling = _placeholder("MinkowskiAvgPooling")
MinkowskiSumPooling = _placeholder("MinkowskiSumPooling")
MinkowskiDropout = _placeholder("MinkowskiDropout")
MinkowskiGELU = _placeholder("MinkowskiGELU")
MinkowskiSinusoidal = _placeholder("MinkowskiSinusoidal")
MinkowskiToSparseTensor = _placeholder("MinkowskiToSparseTensor")
MinkowskiPoolingTranspose = _placeholder("MinkowskiPoolingTranspose")
MinkowskiBroadcastMultiplication = _placeholder("MinkowskiBroadcastMultiplication")
MinkowskiTanh = _placeholder("MinkowskiTanh")
MinkowskiSoftmax = _placeholder("MinkowskiSoftmax")
MinkowskiELU = _placeholder("MinkowskiELU")
MinkowskiPReLU = _placeholder("MinkowskiPReLU")
