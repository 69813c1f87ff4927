"""HyperSeg v0.1 (HyperSeg-L PASCAL VOC) on the B200 kernels.

Mirror of the reference's hyperseg/models/hyperseg_v0_1.py.  The older design differs from v1.0 in three ways
that matter to the kernels:
  * the weight mapper ends in a multi-headed grouped 1x1 conv (``Conv2dMulti``, reference :336-362) and returns a
    *list* of weight maps, one per decoder level -- here each head is one hsb_signal2weights_fwd launch that emits
    only the channels the level uses, patch-major;
  * decoder blocks are chains ``MetaPatchConv2d -> BatchNorm2d -> ReLU6`` (reference :205-237): the map is folded
    back and normalised between the pointwise / depthwise / pointwise-linear stages, so the depthwise halo comes
    from neighbouring patches' activations.  Each stage is one hsb_patch_conv*_fwd launch with BN+activation fused;
  * the decoder has one level per pyramid entry including the full-resolution image, and no final upsample.
"""
import numbers
from functools import partial
from itertools import groupby

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import ops
from .hyperseg_v1_0 import HyperGen as _HyperGenV10
from .hyperseg_v1_0 import assemble_level_input, next_multiply
from .meta_patch import MetaPatchConv2d, make_meta_patch_conv2d_block
from .meta_sequential import MetaSequential


class HyperPatchInvertedResidual(nn.Module):
    """pw (if expand != 1) -> dw 3x3 -> pw-linear, each a patch-wise conv block (reference :205-237)."""

    accepts_strided_weights = True

    def __init__(self, in_nc, out_nc, kernel_size=3, stride=1, expand_ratio=1, norm_layer=nn.BatchNorm2d,
                 act_layer=nn.ReLU6(inplace=True), padding_mode='reflect'):
        super().__init__()
        assert stride in [1, 2]
        self.stride = stride
        hidden = int(round(in_nc * expand_ratio))
        self.use_res_connect = self.stride == 1 and in_nc == out_nc
        stages = []
        if expand_ratio != 1:
            stages.append(make_meta_patch_conv2d_block(in_nc, hidden, 1, norm_layer=norm_layer, act_layer=act_layer))
        stages.append(make_meta_patch_conv2d_block(hidden, hidden, kernel_size, stride=stride, groups=hidden,
                                                   norm_layer=norm_layer, act_layer=act_layer,
                                                   padding_mode=padding_mode))
        stages.append(make_meta_patch_conv2d_block(hidden, out_nc, 1, stride=stride, norm_layer=norm_layer,
                                                   act_layer=None))
        self.conv = MetaSequential(*stages)

    @property
    def hyper_params(self):
        return self.conv.hyper_params

    def forward(self, x, w):
        y = self.conv(x, w)
        return x + y if self.use_res_connect else y


def get_image_coordinates(b, h, w, device):
    xs = torch.linspace(-1, 1, steps=w, device=device)
    ys = torch.linspace(-1, 1, steps=h, device=device)
    grid = torch.stack((xs.view(1, w).expand(h, w), ys.view(h, 1).expand(h, w)), dim=0)
    return grid.unsqueeze(0).expand(b, -1, -1, -1)


def divide_feature_legacy(in_feature, out_features, min_unit=8):
    """The older channel split used by Conv2dMulti (reference :366-406): consumer classes, in decreasing order of
    total demand, get floor(proportional share) (at least one unit, rounded down to a multiple of the class size);
    the last class takes what is left."""
    assert in_feature % min_unit == 0, f'in_feature ({in_feature}) must be divisible by min_unit ({min_unit})'
    units = in_feature // min_unit
    order = np.argsort(out_features)
    ranked = np.array(out_features)[order]
    classes = [(size, order[list(members)]) for size, members in groupby(range(len(order)), lambda i: ranked[i])]
    classes.sort(key=lambda c: c[0] * len(c[1]), reverse=True)
    units_per_weight = float(units) / sum(out_features)
    spare = units
    class_units = []
    for ci, (size, members) in enumerate(classes):
        if ci == len(classes) - 1:
            class_units.append(spare)
        else:
            n = len(members)
            share = max(size * n * units_per_weight, 1) // n * n
            class_units.append(share)
            spare -= share
    result = np.zeros(len(out_features), dtype=int)
    for cu, (_, members) in zip(class_units, classes):
        for m in members:
            result[m] = cu // len(members) * min_unit
    return result


class Conv2dMulti(nn.Module):
    """Several grouped convolutions, each reading its own slice of the input channels (reference :336-362)."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1,
                 bias=True, padding_mode='zeros', min_unit=8):
        super().__init__()
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.bias = bias
        self._ranges = [0]
        parts = divide_feature_legacy(in_channels, out_channels, min_unit)
        for i, out_nc in enumerate(out_channels):
            self._ranges.append(self._ranges[-1] + parts[i])
            self.add_module(f'conv_{i}', nn.Conv2d(parts[i], out_nc, kernel_size, stride, padding, dilation, groups,
                                                   bias, padding_mode))

    def head(self, x, i, limit=None):
        """Output of head i; only the first `limit` channels are computed when given."""
        conv = getattr(self, f'conv_{i}')
        if conv.kernel_size != (1, 1) or conv.stride != (1, 1):
            raise NotImplementedError("Conv2dMulti heads are 1x1 / stride 1 in every reference model")
        hp = conv.out_channels if limit is None else int(limit)
        y = ops.signal2weights(x, conv.weight, int(self._ranges[i]), conv.in_channels, hp, conv.groups)
        if conv.bias is not None:
            y = y + conv.bias[:hp].to(y.dtype).view(1, -1, 1, 1)
        return y

    def forward(self, x):
        return [self.head(x, i) for i in range(len(self.out_channels))]

    def extra_repr(self):
        return f'in_channels={self.in_channels}, out_channels={self.out_channels}, bias={self.bias}'


class WeightMapper(nn.Module):
    """v0.1 context head (reference :249-329): stride-2 down path, nearest-neighbour up path with 1x1 "flat" merges,
    then the multi-headed output conv.  Everything but the output conv is stock PyTorch."""

    def __init__(self, in_channels, out_channels, levels=2, bias=False, min_unit=8, down_groups=1, flat_groups=1,
                 weight_groups=1, avg_pool=False):
        super().__init__()
        assert levels > 0, 'levels must be greater than zero'
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.levels = levels
        self.bias = bias
        self.avg_pool = avg_pool
        self.down_groups = down_groups
        self.flat_groups = flat_groups
        self.weight_groups = weight_groups
        min_unit = max(min_unit, weight_groups)
        for level in range(levels - 1):
            self.add_module(f'down_{level}', nn.Sequential(
                nn.Conv2d(in_channels, in_channels, kernel_size=2, stride=2, bias=bias, groups=down_groups),
                nn.BatchNorm2d(in_channels), nn.ReLU(inplace=True)))
            self.add_module(f'up_{level}', nn.UpsamplingNearest2d(scale_factor=2))
            flat = [nn.Conv2d(in_channels * 2, in_channels, kernel_size=1, bias=bias, groups=flat_groups),
                    nn.BatchNorm2d(in_channels)]
            if level > 0:
                flat.append(nn.ReLU(inplace=True))
            self.add_module(f'flat_{level}', nn.Sequential(*flat))
        padded = [next_multiply(c, weight_groups) for c in out_channels]
        self.out_conv = Conv2dMulti(in_channels, padded, 1, bias=bias, min_unit=min_unit, groups=weight_groups)

    def _heads(self, x):
        n = len(self.out_channels)
        if self.weight_groups > 1:       # reference :323-324 slices the group padding off afterwards
            return [self.out_conv.head(x, i, self.out_channels[i]) for i in range(n)]
        return [self.out_conv.head(x, i) for i in range(n)]

    def forward(self, x):
        if self.levels <= 1:
            return self._heads(x)
        pyramid = [x]
        for level in range(self.levels - 1):
            pyramid.append(getattr(self, f'down_{level}')(pyramid[-1]))
        if self.avg_pool and tuple(pyramid[-1].shape[-2:]) != (1, 1):
            res = pyramid[-1].shape[-2:]
            pyramid[-1] = F.interpolate(F.adaptive_avg_pool2d(pyramid[-1], 1), res, mode='nearest')
        for level in reversed(range(self.levels - 1)):
            up = getattr(self, f'up_{level}')(pyramid.pop())
            pyramid[-1] = getattr(self, f'flat_{level}')(torch.cat((pyramid[-1], up), dim=1))
        return self._heads(pyramid[-1])

    def extra_repr(self):
        return f'in_channels={self.in_channels}, out_channels={self.out_channels}, bias={self.bias}'


class MultiScaleDecoder(nn.Module):
    def __init__(self, feat_channels, in_nc=3, num_classes=3, kernel_sizes=3, level_layers=1, norm_layer=nn.BatchNorm2d,
                 act_layer=nn.ReLU6(inplace=True), out_kernel_size=1, expand_ratio=1, with_out_fc=False, dropout=None):
        super().__init__()
        n = len(feat_channels)
        if isinstance(kernel_sizes, numbers.Number):
            kernel_sizes = (kernel_sizes,) * n
        if isinstance(level_layers, numbers.Number):
            level_layers = (level_layers,) * n
        assert len(kernel_sizes) == n, f'kernel_sizes ({len(kernel_sizes)}) must be of size {n}'
        assert len(level_layers) == n, f'level_layers ({len(level_layers)}) must be of size {n}'
        self.level_layers = level_layers
        self.levels = len(level_layers)
        self.layer_params = []
        feat_channels = feat_channels[::-1]

        carried = 0
        for level in range(self.levels):
            width = feat_channels[level]
            carried += width
            blocks = []
            for layer in range(level_layers[level]):
                if not with_out_fc and level == self.levels - 1 and layer == level_layers[level] - 1:
                    width = num_classes
                if kernel_sizes[level] > 1:
                    blocks.append(HyperPatchInvertedResidual(carried + 2, width, kernel_sizes[level],
                                                             expand_ratio=expand_ratio, norm_layer=norm_layer,
                                                             act_layer=act_layer))
                else:
                    blocks.append(make_meta_patch_conv2d_block(carried + 2, width, kernel_sizes[level]))
                carried = width
            self.add_module(f'level_{level}', MetaSequential(*blocks))

        if with_out_fc:
            tail = [nn.Dropout2d(dropout, True)] if dropout is not None else []
            tail.append(MetaPatchConv2d(carried, num_classes, out_kernel_size, padding=out_kernel_size // 2))
            self.out_fc = MetaSequential(*tail)
        else:
            self.out_fc = None

        self.hyper_params = 0
        self._ranges = [0]
        self.param_groups = []
        for level in range(self.levels):
            hp = getattr(self, f'level_{level}').hyper_params
            self.hyper_params += hp
            self._ranges.append(self.hyper_params)
            self.param_groups.append(hp)
        if with_out_fc:
            self.hyper_params += self.out_fc.hyper_params
            self.param_groups.append(self.out_fc.hyper_params)
        self._ranges.append(self.hyper_params)

    def forward(self, x, w):
        assert isinstance(w, (list, tuple))
        assert len(x) <= self.levels
        p = None
        for level in range(len(x)):
            skip = x[-level - 1]
            coords = get_image_coordinates(1, *skip.shape[-2:], skip.device)
            p = assemble_level_input(coords, skip, p)
            p = getattr(self, f'level_{level}')(p, w[level])
        if self.out_fc is not None:
            p = self.out_fc(p, w[-1])
        return p


class HyperGen(_HyperGenV10):
    def __init__(self, backbone, weight_mapper, in_nc=3, num_classes=3, kernel_sizes=3, level_layers=1, expand_ratio=1,
                 groups=1, inference_hflip=False, inference_gather='mean', with_out_fc=False, decoder_dropout=None):
        nn.Module.__init__(self)
        self.inference_hflip = inference_hflip
        self.inference_gather = inference_gather
        self.backbone = backbone()
        feat_channels = [in_nc] + self.backbone.feat_channels[:-1]
        self.decoder = MultiScaleDecoder(feat_channels, 3, num_classes, kernel_sizes, level_layers,
                                         with_out_fc=with_out_fc, out_kernel_size=1, expand_ratio=expand_ratio,
                                         dropout=decoder_dropout)
        self.weight_mapper = weight_mapper(self.backbone.feat_channels[-1], self.decoder.param_groups)


def hyperseg_efficientnet(model_name, pretrained=False, levels=3, down_groups=1, flat_groups=1, weight_groups=1,
                          avg_pool=True, weights_path=None, **kwargs):
    from .efficientnet import efficientnet
    weight_mapper = partial(WeightMapper, levels=levels, down_groups=down_groups, flat_groups=flat_groups,
                            weight_groups=weight_groups, avg_pool=avg_pool)
    backbone = partial(efficientnet, model_name, pretrained=pretrained, head=None, return_features=True)
    model = HyperGen(backbone, weight_mapper, **kwargs)
    if weights_path is not None:
        checkpoint = torch.load(weights_path, map_location='cpu')
        model.load_state_dict(checkpoint['state_dict'], strict=True)
    return model
