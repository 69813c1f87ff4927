"""HyperSeg v1.0 "unify" (HyperSeg-S Cityscapes) on the B200 kernels.

Mirror of the reference's hyperseg/models/hyperseg_v1_0_unify.py.  It differs from hyperseg_v1_0 only in the
weight plumbing: the signal->weights heads are hoisted out of the dynamic layers into ``WeightLayer`` modules
owned by the decoder (reference :287-309), the dynamic layers take ready-made weights, and from level
``unify_level - 1`` on ONE head emits the weights of all remaining levels, each level taking its channel range
(reference :242-249).  Here a WeightLayer is one hsb_signal2weights_fwd launch writing patch-major rows; the
per-level range is passed to the patch kernels as a strided view (no copy).

state_dict keys: ``decoder.weight_blocks.{i}.signal2weights.weight``, ``decoder.level_blocks.{i}...``.
"""
import numbers
from functools import partial

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import ops
from .hyperseg_v1_0 import (assemble_level_input, HyperPatch, HyperPatchConv2d, HyperPatchInvertedResidual, HyperPatchNoPadding,  # noqa: F401
                            WeightMapper, divide_feature, make_hyper_patch_conv2d_block, next_multiply)
from .hyperseg_v1_0 import HyperGen as _HyperGenV10
from .hyperseg_v1_0 import MultiScaleDecoder as _DecoderV10
from .meta_sequential import MetaSequential


class WeightLayer(nn.Module):
    """A signal->weights head detached from the layer that consumes the weights (reference :287-309)."""

    def __init__(self, target_params):
        super().__init__()
        self.target_params = target_params
        self.signal_channels = None
        self.signal_index = None
        self.signal2weights = None

    def init_signal2weights(self, signal_channels, signal_index=0, groups=1):
        self.signal_channels = signal_channels
        self.signal_index = signal_index
        self.signal2weights = nn.Conv2d(signal_channels, next_multiply(self.target_params, groups), 1, bias=False,
                                        groups=groups)

    def apply_signal2weights(self, s):
        if self.signal2weights is None:
            return s
        head = self.signal2weights
        return ops.signal2weights(s, head.weight, int(self.signal_index), int(self.signal_channels),
                                  int(self.target_params), head.groups)

    def forward(self, s):
        return self.apply_signal2weights(s)


def get_hyper_params(model):
    counts = []
    for child in model.children():
        if isinstance(child, WeightLayer):
            counts.append(child.target_params)
        else:
            counts += get_hyper_params(child)
    return counts


def init_signal2weights(model, signal_features, signal_index=0, weight_groups=1):
    """Same traversal rule as hyperseg_v1_0.init_signal2weights (reference :273-284); the WeightLayers are siblings
    in one ModuleList, so here the signal offset does advance from head to head."""
    for child in model.children():
        if isinstance(child, WeightLayer):
            channels = signal_features.pop(0)
            groups = weight_groups.pop(0) if isinstance(weight_groups, list) else weight_groups
            child.init_signal2weights(channels, signal_index, groups)
            signal_index += channels
        else:
            init_signal2weights(child, signal_features, signal_index, weight_groups)


class MultiScaleDecoder(nn.Module):
    def __init__(self, feat_channels, signal_channels, num_classes=3, kernel_sizes=3, level_layers=1,
                 level_channels=None, norm_layer=nn.BatchNorm2d, act_layer=nn.ReLU6(inplace=True), out_kernel_size=1,
                 expand_ratio=1, groups=1, weight_groups=1, with_out_fc=False, dropout=None,
                 coords_res=None, unify_level=None):
        super().__init__()
        n_levels = len(level_channels)
        per_level = lambda v: (v,) * n_levels if isinstance(v, numbers.Number) else v
        kernel_sizes, level_layers, expand_ratio = per_level(kernel_sizes), per_level(level_layers), per_level(expand_ratio)
        assert len(kernel_sizes) == n_levels, f'kernel_sizes ({len(kernel_sizes)}) must be of size {n_levels}'
        assert len(level_layers) == n_levels, f'level_layers ({len(level_layers)}) must be of size {n_levels}'
        assert len(expand_ratio) == n_levels, f'expand_ratio ({len(expand_ratio)}) must be of size {n_levels}'
        if isinstance(groups, (list, tuple)):
            assert len(groups) == n_levels, f'groups ({len(groups)}) must be of size {n_levels}'
        self.level_layers = level_layers
        self.levels = n_levels
        self.unify_level = unify_level
        self.layer_params = []
        self.coords_cache = {}
        self.weight_groups = weight_groups
        self.level_blocks = nn.ModuleList()
        self.weight_blocks = nn.ModuleList()
        self._ranges = [0]
        feat_channels = feat_channels[::-1]

        carried = 0
        for level in range(n_levels):
            width = feat_channels[level] if level_channels is None else level_channels[level]
            carried += feat_channels[level]
            blocks = []
            for layer in range(level_layers[level]):
                last = level == n_levels - 1 and layer == level_layers[level] - 1
                if last and not with_out_fc:
                    width = num_classes
                if kernel_sizes[level] > 1:
                    blocks.append(HyperPatchInvertedResidual(carried + 2, width, kernel_sizes[level],
                                                             expand_ratio=expand_ratio[level],
                                                             norm_layer=norm_layer, act_layer=act_layer))
                else:
                    g = groups[level] if isinstance(groups, (list, tuple)) else groups
                    blocks.append(make_hyper_patch_conv2d_block(carried + 2, width, kernel_sizes[level], groups=g))
                carried = width
            self.level_blocks.append(MetaSequential(*blocks))
            if level < unify_level - 1:
                self.weight_blocks.append(WeightLayer(self.level_blocks[-1].hyper_params))
            else:
                self._ranges.append(self._ranges[-1] + self.level_blocks[-1].hyper_params)
                if level == n_levels - 1:
                    shared = sum(b.hyper_params for b in self.level_blocks[unify_level - 1:])
                    self.weight_blocks.append(WeightLayer(shared))

        if with_out_fc:
            tail = [nn.Dropout2d(dropout, True)] if dropout is not None else []
            tail.append(HyperPatchConv2d(carried, num_classes, out_kernel_size, padding=out_kernel_size // 2))
            self.out_fc = MetaSequential(*tail)
        else:
            self.out_fc = None

        if coords_res is not None:
            for res in coords_res:
                for i in range(n_levels):
                    h, w = res[0] // 2 ** i, res[1] // 2 ** i
                    self.register_buffer(f'coord{h}_{w}', self.cache_image_coordinates(h, w))

        self.param_groups = get_hyper_params(self)
        min_unit = max(weight_groups)
        signal_features = divide_feature(signal_channels, self.param_groups, min_unit=min_unit)
        head_groups = list(weight_groups) if isinstance(weight_groups, list) else weight_groups
        init_signal2weights(self, list(signal_features), weight_groups=head_groups)
        self.hyper_params = sum(self.param_groups)

    cache_image_coordinates = _DecoderV10.cache_image_coordinates
    _coordinate_grid = staticmethod(_DecoderV10._coordinate_grid)
    get_image_coordinates = _DecoderV10.get_image_coordinates

    @staticmethod
    def _single_ir(block):
        """The inverted-residual layer of a level that consists of nothing else (every shipped unify configuration)."""
        if isinstance(block, HyperPatchInvertedResidual):
            return block
        if isinstance(block, MetaSequential) and len(block) == 1 and isinstance(block[0], HyperPatchInvertedResidual):
            return block[0]
        return None

    def forward_features(self, x, s):
        p = None
        w = None
        for level in range(self.levels):
            skip = x[-level - 1]
            coords = self.get_image_coordinates(1, *skip.shape[-2:], skip.device)
            p = assemble_level_input(coords, skip, p)
            block = self.level_blocks[level]
            if level < self.unify_level - 1:
                wl = self.weight_blocks[level]
                y = None
                ir = self._single_ir(block)
                if ir is not None and wl.signal2weights is not None:
                    y = ir.forward_arranged(p, s, wl.signal2weights, wl.signal_index, wl.signal_channels)
                p = block(p, wl(s)) if y is None else y
            else:
                shared = self.weight_blocks[self.unify_level - 1]       # one head for all remaining levels
                i = level - self.unify_level + 1
                y = None
                ir = self._single_ir(block)
                if ir is not None and shared.signal2weights is not None:
                    # the head writes this level's slice straight in the fused kernel's operand order
                    y = ir.forward_arranged(p, s, shared.signal2weights, shared.signal_index, shared.signal_channels,
                                            hp_offset=int(self._ranges[i]))
                if y is None:
                    if w is None:
                        w = shared(s)
                    y = block(p, w[:, self._ranges[i]:self._ranges[i + 1]])
                p = y
        if self.out_fc is not None:
            p = self.out_fc(p, s)
        return p

    def forward(self, x, s):
        p = self.forward_features(x, s)
        if p.shape[2:] != x[0].shape[2:]:
            p = F.interpolate(p, x[0].shape[2:], mode='bilinear', align_corners=False)
        return p


class HyperGen(_HyperGenV10):
    def __init__(self, backbone, weight_mapper, in_nc=3, num_classes=3, kernel_sizes=3, level_layers=1,
                 level_channels=None, expand_ratio=1, groups=1, weight_groups=1, inference_hflip=False,
                 inference_gather='mean', with_out_fc=False, decoder_groups=1, decoder_dropout=None, coords_res=None,
                 unify_level=None):
        nn.Module.__init__(self)
        self.inference_hflip = inference_hflip
        self.inference_gather = inference_gather
        self.backbone = backbone()
        feat_channels = [in_nc] + self.backbone.feat_channels[:-1]
        self.decoder = MultiScaleDecoder(feat_channels, self.backbone.feat_channels[-1], num_classes, kernel_sizes,
                                         level_layers, level_channels, with_out_fc=with_out_fc, out_kernel_size=1,
                                         expand_ratio=expand_ratio, groups=decoder_groups, weight_groups=weight_groups,
                                         dropout=decoder_dropout, coords_res=coords_res, unify_level=unify_level)
        self.weight_mapper = weight_mapper(self.backbone.feat_channels[-1], self.decoder.param_groups)


def hyperseg_efficientnet(model_name, pretrained=False, out_feat_scale=0.25, levels=3, weights_path=None, **kwargs):
    from .efficientnet import efficientnet
    weight_mapper = partial(WeightMapper, levels=levels)
    backbone = partial(efficientnet, model_name, pretrained=pretrained, out_feat_scale=out_feat_scale, head=None,
                       return_features=True)
    model = HyperGen(backbone, weight_mapper, **kwargs)
    if weights_path is not None:
        checkpoint = torch.load(weights_path, map_location='cpu')
        model.load_state_dict(checkpoint['state_dict'], strict=True)
    return model
