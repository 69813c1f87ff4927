"""Per-sample dynamic convolution.

Mirror of the reference's ``MetaConv2d`` (hyperseg/models/layers/meta_conv.py:141-186): every batch element
is convolved with its own weight set ``w[n]`` of ``hyper_params`` values, laid out as
(out_channels, in_channels/groups, kh, kw).  The reference folds the batch into the channel axis and calls
``F.conv2d(groups=N*groups)``; here one CUDA kernel indexes the per-sample weights directly.
"""
import numpy as np
import torch.nn as nn
from torch.nn.modules.utils import _pair

from .. import ops
from .meta_sequential import MetaSequential

_PADDING_MODES = ("zeros", "reflect", "replicate", "circular")


class MetaConv2d(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1,
                 padding_mode='zeros'):
        super().__init__()
        if in_channels % groups != 0:
            raise ValueError('in_channels must be divisible by groups')
        if out_channels % groups != 0:
            raise ValueError('out_channels must be divisible by groups')
        if padding_mode not in _PADDING_MODES:
            raise ValueError(f"padding_mode must be one of {set(_PADDING_MODES)}, but got padding_mode='{padding_mode}'")
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.kernel_size = _pair(kernel_size)
        self.stride = _pair(stride)
        self.padding = _pair(padding)
        self.dilation = _pair(dilation)
        self.groups = groups
        self.padding_mode = padding_mode
        self._padding_repeated_twice = self.padding + self.padding
        self.hyper_params = np.prod((out_channels, in_channels // groups) + self.kernel_size)

    def forward(self, x, w):
        assert x.shape[0] == w.shape[0]
        if self.stride != (1, 1):
            raise NotImplementedError("hyperseg_b200 MetaConv2d kernels cover stride 1 (the only stride the "
                                      "reference models use)")
        return ops.meta_conv2d(x, w, self.out_channels, self.kernel_size, self.padding, self.dilation,
                               self.groups, self.padding_mode)

    def extra_repr(self):
        parts = [f'{self.in_channels}, {self.out_channels}, kernel_size={self.kernel_size}, stride={self.stride}']
        if any(self.padding):
            parts.append(f'padding={self.padding}')
        if self.dilation != (1, 1):
            parts.append(f'dilation={self.dilation}')
        if self.groups != 1:
            parts.append(f'groups={self.groups}')
        if self.padding_mode != 'zeros':
            parts.append(f'padding_mode={self.padding_mode}')
        return ', '.join(parts)


def make_meta_conv2d_block(in_nc, out_nc, kernel_size=3, stride=1, padding=None, dilation=1, groups=1,
                           padding_mode='reflect', norm_layer=nn.BatchNorm2d, act_layer=nn.ReLU(True), dropout=None):
    """MetaConv2d [+ norm] [+ activation] [+ dropout] chained in a MetaSequential (meta_conv.py:202-230)."""
    assert dropout is None or isinstance(dropout, float)
    padding = kernel_size // 2 if padding is None else padding
    layers = [MetaConv2d(in_nc, out_nc, kernel_size, stride, padding, dilation, groups, padding_mode)]
    if norm_layer is not None:
        layers.append(norm_layer(out_nc))
    if act_layer is not None:
        layers.append(act_layer)
    if dropout is not None:
        layers.append(nn.Dropout(dropout))
    return MetaSequential(*layers)
