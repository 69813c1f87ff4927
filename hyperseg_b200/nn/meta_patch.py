"""Patch-wise dynamic convolution taking ready-made weights.

Mirror of the reference's ``MetaPatch`` / ``MetaPatchConv2d`` (hyperseg/models/layers/meta_patch.py:19-57,
:190-225), used by HyperSeg-L (hyperseg_v0_1).  ``forward(x, weight)``: ``weight`` is (B, hyper_params, fh, fw);
the map is tiled into fh x fw patches, each patch is padded from the *whole padded map* (so interior halos
are neighbouring pixels) and convolved with its own weights.  The reference does pad -> F.unfold ->
MetaConv2d(groups = tiles) -> F.fold; here it is one kernel, with eval BatchNorm + activation fused when
the enclosing MetaSequential offers them.
"""
import torch.nn as nn
from torch.nn.modules.utils import _pair

from .. import ops
from .meta_conv import MetaConv2d, _PADDING_MODES
from .meta_sequential import MetaSequential


def _run_patch_conv(conv, padding, padding_mode, x, weight, fused_norm, fused_act):
    if conv.stride != (1, 1):
        raise NotImplementedError("patch-wise convolution kernels cover stride 1 only")
    scale = shift = None
    if fused_norm is not None:
        scale, shift = ops.fold_bn(fused_norm)
    kh, kw = conv.kernel_size
    if (kh, kw) == (1, 1) and padding == (0, 0):
        return ops.patch_conv1x1(x, weight, conv.out_channels, conv.groups, scale, shift, fused_act)
    return ops.patch_conv(x, weight, conv.out_channels, conv.kernel_size, padding, conv.dilation, conv.groups,
                          padding_mode, scale, shift, fused_act)


class MetaPatch(nn.Module):
    supports_fused_epilogue = True
    accepts_strided_weights = True

    def __init__(self, module: nn.Module, padding=0, padding_mode='reflect'):
        super().__init__()
        if padding_mode not in _PADDING_MODES:
            raise ValueError(f"padding_mode must be one of {set(_PADDING_MODES)}, but got padding_mode='{padding_mode}'")
        if not isinstance(module, MetaConv2d):
            raise NotImplementedError("hyperseg_b200 MetaPatch wraps MetaConv2d (the only module the reference "
                                      "ever wraps)")
        self.hyper_module = module
        self.padding = _pair(padding)
        self.padding_mode = padding_mode
        self._padding_repeated_twice = self.padding + self.padding

    @property
    def hyper_params(self):
        return self.hyper_module.hyper_params

    def forward(self, x, weight, fused_norm=None, fused_act=None):
        return _run_patch_conv(self.hyper_module, self.padding, self.padding_mode, x, weight, fused_norm, fused_act)


class MetaPatchConv2d(MetaPatch):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1,
                 padding_mode='reflect'):
        super().__init__(MetaConv2d(in_channels, out_channels, kernel_size, stride, 0, dilation, groups),
                         padding, padding_mode)

    in_channels = property(lambda self: self.hyper_module.in_channels)
    out_channels = property(lambda self: self.hyper_module.out_channels)
    kernel_size = property(lambda self: self.hyper_module.kernel_size)
    groups = property(lambda self: self.hyper_module.groups)

    def __repr__(self):
        m = self.hyper_module
        parts = [f'{m.in_channels}, {m.out_channels}, kernel_size={m.kernel_size}, stride={m.stride}']
        if any(self.padding):
            parts.append(f'padding={self.padding}')
        if m.dilation != (1, 1):
            parts.append(f'dilation={m.dilation}')
        if m.groups != 1:
            parts.append(f'groups={m.groups}')
        if self.padding_mode != 'zeros':
            parts.append(f'padding_mode={self.padding_mode}')
        return f"{self.__class__.__name__}({', '.join(parts)})"


def make_meta_patch_conv2d_block(in_nc, out_nc, kernel_size=3, stride=1, padding=None, dilation=1, groups=1,
                                 padding_mode='reflect', norm_layer=nn.BatchNorm2d, act_layer=nn.ReLU(True),
                                 dropout=None):
    """MetaPatchConv2d [+ norm] [+ activation] [+ dropout] in a MetaSequential (meta_patch.py:228-257)."""
    assert dropout is None or isinstance(dropout, float)
    padding = kernel_size // 2 if padding is None else padding
    layers = [MetaPatchConv2d(in_nc, out_nc, kernel_size, stride, padding, dilation, groups, padding_mode)]
    if norm_layer is not None:
        layers.append(norm_layer(out_nc))
    if act_layer is not None:
        layers.append(act_layer)
    if dropout is not None:
        layers.append(nn.Dropout(dropout))
    return MetaSequential(*layers)
