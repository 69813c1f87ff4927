"""EfficientNet feature pyramid used as HyperSeg's encoder -- stock PyTorch, not part of the hot path.

Written to be parameter-compatible with the reference's backbone (hyperseg/models/backbones/efficientnet.py,
a fork of lukemelas/EfficientNet-PyTorch): identical module names (``_conv_stem``, ``_bn0``, ``_blocks.N.
_expand_conv`` ... ``_feat_fc_i``, ``_conv_head``, ``_bn1``), so reference checkpoints load strictly, and
identical arithmetic, including two properties of the reference that matter for parity:

  * TensorFlow-"SAME" padding is *static*: the asymmetric pad of every strided convolution is derived once
    from the network's nominal resolution (240 for B1, 300 for B3; efficientnet_utils.py:247-274), not from
    the actual input size;
  * the feature returned for each resolution is the output of the last block before the next stride-2
    stage, optionally reduced by a 1x1 conv + BN (``out_feat_scale``, efficientnet.py:207-222,319-363).
"""
import math
from collections import namedtuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import ops

# (repeats, kernel, stride, expand, in, out, se_ratio) of the EfficientNet-B0 stages
_B0_STAGES = (
    (1, 3, 1, 1, 32, 16, 0.25),
    (2, 3, 2, 6, 16, 24, 0.25),
    (2, 5, 2, 6, 24, 40, 0.25),
    (3, 3, 2, 6, 40, 80, 0.25),
    (3, 5, 1, 6, 80, 112, 0.25),
    (4, 5, 2, 6, 112, 192, 0.25),
    (1, 3, 1, 6, 192, 320, 0.25),
)
# name -> (width multiplier, depth multiplier, nominal resolution, dropout)
_COMPOUND = {
    'efficientnet-b0': (1.0, 1.0, 224, 0.2), 'efficientnet-b1': (1.0, 1.1, 240, 0.2),
    'efficientnet-b2': (1.1, 1.2, 260, 0.3), 'efficientnet-b3': (1.2, 1.4, 300, 0.3),
    'efficientnet-b4': (1.4, 1.8, 380, 0.4), 'efficientnet-b5': (1.6, 2.2, 456, 0.4),
    'efficientnet-b6': (1.8, 2.6, 528, 0.5), 'efficientnet-b7': (2.0, 3.1, 600, 0.5),
    'efficientnet-b8': (2.2, 3.6, 672, 0.5), 'efficientnet-l2': (4.3, 5.3, 800, 0.5),
}
_BN_MOMENTUM, _BN_EPS = 0.01, 1e-3     # TF momentum 0.99 in PyTorch's convention
_DIVISOR = 8

Stage = namedtuple('Stage', 'repeats kernel stride expand cin cout se_ratio')


def _scale_width(channels, width_mult):
    if not width_mult:
        return channels
    scaled = channels * width_mult
    rounded = max(_DIVISOR, int(scaled + _DIVISOR / 2) // _DIVISOR * _DIVISOR)
    if rounded < 0.9 * scaled:
        rounded += _DIVISOR
    return int(rounded)


def _scale_depth(repeats, depth_mult):
    return repeats if not depth_mult else int(math.ceil(depth_mult * repeats))


def _out_size(size, stride):
    return [int(math.ceil(size[0] / stride)), int(math.ceil(size[1] / stride))]


class SamePadConv2d(nn.Conv2d):
    """Conv2d with TF 'SAME' padding frozen at construction for a nominal input size."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, image_size=None, **kwargs):
        super().__init__(in_channels, out_channels, kernel_size, stride, **kwargs)
        ih, iw = image_size
        kh, kw = self.weight.shape[-2:]
        sh, sw = self.stride
        need_h = max((math.ceil(ih / sh) - 1) * sh + (kh - 1) * self.dilation[0] + 1 - ih, 0)
        need_w = max((math.ceil(iw / sw) - 1) * sw + (kw - 1) * self.dilation[1] + 1 - iw, 0)
        # (left, right, top, bottom): the odd pixel goes to the right / bottom
        self.same_pad = (need_w // 2, need_w - need_w // 2, need_h // 2, need_h - need_h // 2)
        # a symmetric pad is handed to the convolution itself (no padded copy of the activation)
        self.symmetric = need_w % 2 == 0 and need_h % 2 == 0
        self.conv_pad = (need_h // 2, need_w // 2)

    def forward(self, x):
        if self.symmetric:
            return F.conv2d(x, self.weight, self.bias, self.stride, self.conv_pad, self.dilation, self.groups)
        x = F.pad(x, self.same_pad)
        return F.conv2d(x, self.weight, self.bias, self.stride, self.padding, self.dilation, self.groups)


def _bn(channels):
    return nn.BatchNorm2d(channels, momentum=_BN_MOMENTUM, eps=_BN_EPS)


class MBConvBlock(nn.Module):
    """Mobile inverted bottleneck with squeeze-and-excitation."""

    def __init__(self, cin, cout, kernel, stride, expand, se_ratio, image_size, allow_skip):
        super().__init__()
        mid = cin * expand
        self.expand = expand != 1
        if self.expand:
            self._expand_conv = SamePadConv2d(cin, mid, 1, image_size=image_size, bias=False)
            self._bn0 = _bn(mid)
        self._depthwise_conv = SamePadConv2d(mid, mid, kernel, stride, image_size=image_size, groups=mid, bias=False)
        self._bn1 = _bn(mid)
        image_size = _out_size(image_size, stride)
        self.has_se = se_ratio is not None and 0 < se_ratio <= 1
        if self.has_se:
            squeezed = max(1, int(cin * se_ratio))
            self._se_reduce = SamePadConv2d(mid, squeezed, 1, image_size=(1, 1))
            self._se_expand = SamePadConv2d(squeezed, mid, 1, image_size=(1, 1))
        self._project_conv = SamePadConv2d(mid, cout, 1, image_size=image_size, bias=False)
        self._bn2 = _bn(cout)
        self.skip = allow_skip and stride == 1 and cin == cout

    def forward(self, x, drop_connect_rate=None):
        y = x
        if self.expand:
            y = _norm_act(self._bn0, self._expand_conv(y), True)
        if self.has_se:
            y, squeezed = _norm_act(self._bn1, self._depthwise_conv(y), True, pool=True)
            fused = squeezed.dim() == 3         # engine path: per-chunk sums from the epilogue kernel (N, chunks, C)
            if fused:
                squeezed = ops.pooled_mean(squeezed, y.shape[2] * y.shape[3], y.dtype)
            gate = self._se_expand(F.silu(self._se_reduce(squeezed)))
            y = ops.channel_gate_nhwc_(y, gate) if fused else torch.sigmoid(gate) * y
        else:
            y = _norm_act(self._bn1, self._depthwise_conv(y), True)
        y = self._project_conv(y)
        if self.skip and drop_connect_rate and self.training:
            y = self._bn2(y)
            keep = 1 - drop_connect_rate
            mask = torch.floor(keep + _uniform_per_sample(y))
            return y / keep * mask + x
        return _norm_act(self._bn2, y, False, residual=x if self.skip else None)


class FoldedBatchNorm(nn.Module):
    """What an inference engine leaves in a BatchNorm's slot after folding the scale into the convolution before
    it: the per-channel shift.  ``shift32`` (set by the engine once the module sits on its device) switches
    :func:`_norm_act` to the one-pass channels-last epilogue kernel."""

    def __init__(self, shift):
        super().__init__()
        self.register_buffer("shift", shift.detach().clone())
        self.shift_master = shift.detach().float().clone()      # survives .to(dtype=bfloat16)
        self.shift32 = None

    def forward(self, y):
        return y + self.shift.to(y.dtype).view(1, -1, 1, 1)


def _norm_act(bn, y, swish, residual=None, pool=False):
    """BatchNorm (or its folded remainder) -> optional swish -> optional skip add [-> squeeze for SE]: the
    reference's efficientnet.py:97-98, :101-102, :106, :113, :122 -- one kernel on the engine path."""
    shift = getattr(bn, "shift32", None)
    if shift is not None and ops.nhwc_epilogue_ok(y) and (residual is None or ops.nhwc_epilogue_ok(residual)):
        return ops.bias_act_nhwc_(y, shift, "silu" if swish else "none", residual, pool)
    y = bn(y)
    if swish:
        y = F.silu(y)
    if residual is not None:
        y = y + residual
    if pool:
        return y, F.adaptive_avg_pool2d(y, 1)
    return y


def _uniform_per_sample(y):
    """One U[0,1) draw per batch element on y's device (drop-connect mask source; tests swap in a
    CPU-generator version so masks match the reference's CPU run)."""
    return torch.rand([y.shape[0], 1, 1, 1], dtype=y.dtype, device=y.device)


class EfficientNet(nn.Module):
    def __init__(self, model_name, out_feat_scale=0.25, head=None, return_features=True, pool=False,
                 drop_connect_rate=0.2, num_classes=1000):
        super().__init__()
        if model_name not in _COMPOUND:
            raise ValueError('model_name should be one of: ' + ', '.join(_COMPOUND))
        width, depth, res, dropout = _COMPOUND[model_name]
        self.return_features = return_features
        self.pool = pool
        self.drop_connect_rate = drop_connect_rate
        self.out_feat_scale = out_feat_scale
        size = [res, res]

        stem = _scale_width(32, width)
        self._conv_stem = SamePadConv2d(3, stem, 3, 2, image_size=size, bias=False)
        self._bn0 = _bn(stem)
        stem_size = size
        size = _out_size(size, 2)

        self._blocks = nn.ModuleList()
        tap = []            # tap[i]: block i ends a resolution
        widths = []
        for spec in _B0_STAGES:
            st = Stage(*spec)
            cin, cout = _scale_width(st.cin, width), _scale_width(st.cout, width)
            repeats = _scale_depth(st.repeats, depth)
            if st.stride > 1:
                tap[-1] = True
            tap += [False] * repeats
            widths += [cout] * repeats
            self._blocks.append(MBConvBlock(cin, cout, st.kernel, st.stride, st.expand, st.se_ratio, size, False))
            size = _out_size(size, st.stride)
            for _ in range(repeats - 1):
                self._blocks.append(MBConvBlock(cout, cout, st.kernel, 1, st.expand, st.se_ratio, size, True))
        tap[-1] = True
        self._tap = tap
        self.feat_channels = [c for c, t in zip(widths, tap) if t]

        if out_feat_scale is not None:
            for i, cin in enumerate(self.feat_channels):
                scale = out_feat_scale[i] if isinstance(out_feat_scale, (list, tuple)) else out_feat_scale
                cout = int(round(cin * scale))
                if scale != 1.:
                    self.add_module(f'_feat_fc_{i}', nn.Sequential(
                        SamePadConv2d(cin, cout, 1, image_size=stem_size, bias=False), _bn(cout)))
                else:
                    setattr(self, f'_feat_fc_{i}', None)
                self.feat_channels[i] = cout

        head_in = widths[-1]
        head_out = _scale_width(1280, width)
        self.feat_channels.append(head_out)
        self._conv_head = SamePadConv2d(head_in, head_out, 1, image_size=size, bias=False)
        self._bn1 = _bn(head_out)
        self._avg_pooling = nn.AdaptiveAvgPool2d(1)
        self._dropout = nn.Dropout(dropout)
        self._fc = head(head_out, num_classes) if head is not None else None

    def _trunk(self, x, collect):
        x = _norm_act(self._bn0, self._conv_stem(x), True)
        feats = []
        n = len(self._blocks)
        for i, block in enumerate(self._blocks):
            rate = self.drop_connect_rate * float(i) / n if self.drop_connect_rate else None
            x = block(x, rate)
            if collect and self._tap[i]:
                fc = getattr(self, f'_feat_fc_{len(feats)}', None) if self.out_feat_scale is not None else None
                feats.append(x if fc is None else fc(x))
        x = _norm_act(self._bn1, self._conv_head(x), True)
        return x, feats

    def forward(self, x):
        x, feats = self._trunk(x, self.return_features)
        if self.pool:
            x = self._avg_pooling(x).flatten(1)
        x = self._dropout(x)
        if self._fc is not None:
            x = self._fc(x)
        if self.return_features:
            return feats + [x]
        return x


def efficientnet(model_name, pretrained=False, head=nn.Linear, **kwargs):
    """Factory with the reference's call signature (efficientnet.py:493-502)."""
    if pretrained:
        raise RuntimeError("pretrained EfficientNet weights are fetched from the network by the reference; "
                           "load a checkpoint with load_state_dict instead")
    return EfficientNet(model_name, head=head, **kwargs)
