"""HyperSeg v1.0 (HyperSeg-M Cityscapes, HyperSeg-S/L CamVid) on the B200 kernels.

Host-side mirror of the reference's hyperseg/models/hyperseg_v1_0.py: the same classes, constructor
arguments, attributes, forward signatures and state_dict keys, so a checkpoint written by the reference
loads with ``strict=True`` and the classes can stand in for the reference's.  What changes is what the
dynamic layers *do*:

  HyperPatchNoPadding.forward          -> one call of hsb_signal2weights_fwd + hsb_patch_conv1x1_fwd
                                          (reference :486-498 = 2 permute copies + grouped conv + permute copy)
  HyperPatchInvertedResidual.forward   -> hsb_signal2weights_fwd + hsb_patch_ir_fwd
                                          (reference :328-376 = pad, unfold, 3 grouped convs, 3 BN, 2 ReLU6, re-tile)
  HyperPatchConv2d.forward             -> hsb_signal2weights_fwd + hsb_patch_conv_fwd (reference :543-557)

The backbone and the WeightMapper trunk are stock PyTorch, as in the reference; the decoder glue (bilinear upsample
of the previous level + concat with the encoder feature and the coordinate channels, reference :235-240) is one
hsb_decoder_input_fwd launch per level on the GPU.
"""
import numbers
from functools import partial
from itertools import groupby

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.nn.modules.utils import _pair

from .. import ops
from .meta_conv import MetaConv2d, _PADDING_MODES
from .meta_patch import _run_patch_conv
from .meta_sequential import MetaSequential


def next_multiply(x, base):
    """Smallest multiple of ``base`` that is >= x, keeping x's type (reference :451-452)."""
    return type(x)(np.ceil(x / base) * base)


class _SignalHeadMixin:
    """Ownership of a ``signal2weights`` head: grouped 1x1 conv from a slice of the signal map to this
    layer's weights (reference :315-326, :473-484, :531-541).  The nn.Conv2d is kept as the parameter
    container (state_dict key ``signal2weights.weight``); the arithmetic runs in hsb_signal2weights_fwd,
    which only produces the ``hyper_params`` channels that are used and writes them patch-major."""

    _pad_head_to_groups = True

    def _reset_head(self):
        self.signal_channels = None
        self.signal_index = None
        self.signal2weights = None

    def init_signal2weights(self, signal_channels, signal_index=0, groups=1):
        self.signal_channels = signal_channels
        self.signal_index = signal_index
        out = next_multiply(self.hyper_params, groups) if self._pad_head_to_groups else self.hyper_params
        self.signal2weights = nn.Conv2d(signal_channels, out, 1, bias=False, groups=groups)

    def apply_signal2weights(self, s):
        if self.signal2weights is None:
            return s
        head = self.signal2weights
        return ops.signal2weights(s, head.weight, int(self.signal_index), int(self.signal_channels),
                                  int(self.hyper_params), head.groups)


class HyperPatchInvertedResidual(nn.Module, _SignalHeadMixin):
    """Patch-wise inverted residual MetaBlock: pw expand -> BN -> ReLU6 -> dw3x3 -> BN -> ReLU6 -> pw project
    -> BN, evaluated per patch on its 1-pixel halo tile with that patch's own generated weights."""

    accepts_strided_weights = True

    def __init__(self, in_nc, out_nc, kernel_size=3, stride=1, expand_ratio=1, norm_layer=nn.BatchNorm2d,
                 act_layer=nn.ReLU6(inplace=True), padding_mode='reflect'):
        super().__init__()
        assert stride in [1, 2]
        self.stride = stride
        self.padding_mode = padding_mode
        self.padding = (1, 1)
        self._padding_repeated_twice = self.padding + self.padding
        self.in_nc = in_nc
        self.out_nc = out_nc
        self.kernel_size = _pair(kernel_size)
        self.hidden_dim = int(round(in_nc * expand_ratio))
        self.use_res_connect = self.stride == 1 and in_nc == out_nc
        self.act_layer = act_layer
        self.bn1 = norm_layer(self.hidden_dim)
        self.bn2 = norm_layer(self.hidden_dim)
        self.bn3 = norm_layer(self.out_nc)

        # weight vector = [expand (hid x in) | depthwise (hid x kh x kw) | project (out x hid)]
        sizes = (in_nc * self.hidden_dim, np.prod((self.hidden_dim,) + self.kernel_size), self.hidden_dim * out_nc)
        self.hyper_params = 0
        self._ranges = [0]
        for n in sizes:
            self.hyper_params += n
            self._ranges.append(self.hyper_params)
        self._reset_head()

    def _check_supported(self):
        if self.kernel_size != (3, 3) or self.stride != 1 or self.padding_mode != 'reflect':
            raise NotImplementedError("the fused inverted-residual kernel covers the configuration every reference "
                                      "model uses: 3x3 depthwise, stride 1, reflect padding")
        if not isinstance(self.act_layer, nn.ReLU6):
            raise NotImplementedError("the fused inverted-residual kernel hard-wires ReLU6 (the decoder's act_layer)")
        for bn in (self.bn1, self.bn2, self.bn3):
            if not isinstance(bn, nn.BatchNorm2d):
                raise NotImplementedError("only BatchNorm2d norm layers are fused")

    def forward_arranged(self, x, s, head=None, signal_index=None, signal_channels=None, hp_offset=0):
        """Fast inference path: the head writes this block's weights directly in the operand order of the restage-free
        tensor-core kernel (BatchNorm scales folded into the packed head weights), and that kernel consumes them --
        ops.signal2weights_arranged + ops.patch_ir_arranged.  ``head`` is the grouped 1x1 nn.Conv2d that generates the
        weights (default: this layer's own ``signal2weights``); ``hp_offset`` is the first of its output channels that
        belongs to this block (unify: one head feeds several levels).  Returns None when the path does not apply
        (training / autograd, fp32 compute, a shape without an instantiation, residual blocks) -- the caller then takes
        the general path."""
        head = self.signal2weights if head is None else head
        if head is None or self.training or self.use_res_connect or not x.is_cuda or ops._needs_grad(x, s, head.weight):
            return None
        if ops._compute_dtype(x) != torch.bfloat16 or not ops.head_tc_ok(s):
            return None
        (B, _, H, W), (fh, fw) = x.shape, s.shape[-2:]
        if H % fh or W % fw or H // fh != W // fw or W % 8 or (H * W) % 8:
            return None
        if not ops.ir_arranged_supported(self.in_nc, self.hidden_dim, self.out_nc, H // fh):
            return None
        self._check_supported()
        sig_index = int(self.signal_index if signal_index is None else signal_index)
        sig_ch = int(self.signal_channels if signal_channels is None else signal_channels)
        norms = (self.bn1, self.bn2, self.bn3)
        key = (head.weight.data_ptr(), head.weight._version, head.weight.dtype, sig_index, sig_ch, int(hp_offset),
               tuple((t.data_ptr(), t._version) for bn in norms for t in (bn.running_mean, bn.running_var, bn.weight, bn.bias)
                     if t is not None),
               tuple(id(getattr(bn, "_hsb_folded", None)) for bn in norms))
        cached = getattr(self, "_hsb_arranged", None)
        if cached is None or cached[0] != key:
            bns = [ops.fold_bn(bn) for bn in norms]
            bns = [(a.to(x.device, torch.float32).contiguous(), b.to(x.device, torch.float32).contiguous()) for a, b in bns]
            try:
                packed = ops.ArrangedHead(head.weight, sig_index, sig_ch, head.groups, int(hp_offset), self.in_nc, self.hidden_dim,
                                          self.out_nc, bns[0][0], bns[1][0], bns[2][0])
            except ops._lib.HsbError:
                packed = None                              # a tile of this head needs too wide a signal range
            # the packed operand, its tile table and the folded shifts live as long as this module: a CUDA graph captured
            # over this call keeps reading them
            cached = (key, packed, bns)
            self._hsb_arranged = cached
        bns = cached[2]
        if cached[1] is None:
            return None
        w_arr = ops.signal2weights_arranged(s, cached[1])
        xb = x if x.dtype == torch.bfloat16 else x.to(torch.bfloat16)
        return ops.patch_ir_arranged(xb, w_arr, self.hidden_dim, self.out_nc, bns[0][1], bns[1][1], bns[2][1])

    def _run(self, x, s, residual):
        self._check_supported()
        if not residual and self.signal2weights is not None:
            y = self.forward_arranged(x, s)
            if y is not None:
                return y
        weight = self.apply_signal2weights(s)
        if self.training or ops._needs_grad(x, weight):
            y = self._run_stagewise(x, weight)
            return x + y if residual else y
        return ops.patch_ir(x, weight, self.hidden_dim, self.out_nc, ops.fold_bn(self.bn1), ops.fold_bn(self.bn2),
                            ops.fold_bn(self.bn3), residual=residual)

    def _run_stagewise(self, x, weight):
        """Training path of the block (reference :328-368): the fused kernel has no backward and folds eval-mode
        BatchNorms, so under autograd / train mode the three stages run as separate differentiable patch
        convolutions around this module's own BatchNorms.  The halo tiles are laid out side by side as one
        (fh*(ph+2)) x (fw*(pw+2)) map, which keeps the BatchNorm statistics those of the reference's
        (b*fh*fw, C, ph+2, pw+2) view (same multiset of values per channel)."""
        B, C, H, W = x.shape
        fh, fw = weight.shape[-2:]
        ph, pw = H // fh, W // fw
        kh, kw = ph + 2, pw + 2
        hid, r = self.hidden_dim, self._ranges
        tiles = F.pad(x, self._padding_repeated_twice, mode=self.padding_mode).unfold(2, kh, ph).unfold(3, kw, pw)
        h = tiles.permute(0, 1, 2, 4, 3, 5).reshape(B, C, fh * kh, fw * kw)                    # :342
        h = self.act_layer(self.bn1(ops.patch_conv1x1(h, weight[:, r[0]:r[1]], hid)))           # :350-353
        # depthwise on whole halo tiles with zero padding, keeping each tile's interior == the reference's
        # unpadded ("valid") depthwise of :359
        h = ops.patch_conv(h, weight[:, r[1]:r[2]], hid, (3, 3), (1, 1), (1, 1), hid, 'zeros')
        h = h.reshape(B, hid, fh, kh, fw, kw)[:, :, :, 1:-1, :, 1:-1].reshape(B, hid, H, W)
        h = self.act_layer(self.bn2(h))                                                         # :360-361
        return self.bn3(ops.patch_conv1x1(h, weight[:, r[2]:r[3]], self.out_nc))                # :364-366

    def conv(self, x, s):
        return self._run(x, s, residual=False)

    def forward(self, x, s):
        return self._run(x, s, residual=self.use_res_connect)   # the skip add is fused into the kernel


class WeightMapper(nn.Module):
    """Context head: small U-Net over the 1/32-resolution signal (reference :379-448). Stock PyTorch."""

    def __init__(self, in_channels, out_channels, levels=3, bias=False, min_unit=4, weight_groups=1):
        super().__init__()
        assert levels > 0, 'levels must be greater than zero'
        assert in_channels % 2 == 0, 'in_channels must be divisible by 2'
        if isinstance(weight_groups, (list, tuple)):
            assert len(weight_groups) == len(out_channels), \
                f'groups ({len(weight_groups)}) must be of size {len(out_channels)}'
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.levels = levels
        self.bias = bias
        self.weight_groups = weight_groups
        half = in_channels // 2

        def block(cin, k, stride):
            return nn.Sequential(nn.Conv2d(cin, half, kernel_size=k, stride=stride, bias=bias),
                                 nn.BatchNorm2d(half), nn.ReLU(inplace=True))

        self.down_blocks = nn.ModuleList()
        self.up_blocks = nn.ModuleList()
        self.in_conv = block(in_channels, 1, 1)
        for _ in range(levels - 1):
            self.down_blocks.append(block(half, 2, 2))
            self.up_blocks.append(block(in_channels, 1, 1))
        self.upsample = nn.UpsamplingNearest2d(scale_factor=2)

    def forward(self, x):
        pyramid = [self.in_conv(x)]
        for down in self.down_blocks:
            pyramid.append(down(pyramid[-1]))
        # global context: mean of the coarsest map, broadcast back over it
        x = pyramid[-1]
        coarse_res = x.shape[-2:]
        if tuple(coarse_res) != (1, 1):
            x = F.interpolate(F.adaptive_avg_pool2d(x, 1), coarse_res, mode='nearest')
        for level in reversed(range(self.levels - 1)):
            x = self.up_blocks[level](torch.cat((pyramid.pop(), x), dim=1))
            x = self.upsample(x)
        return torch.cat((pyramid.pop(), x), dim=1)


class HyperPatchNoPadding(nn.Module, _SignalHeadMixin):
    """Patch-wise convolution without halo (used with kernel_size 1): y_patch = W_patch (*) x_patch."""

    supports_fused_epilogue = True
    accepts_strided_weights = True

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, dilation=1, groups=1):
        super().__init__()
        if in_channels % groups != 0:
            raise ValueError('in_channels must be divisible by groups')
        if out_channels % groups != 0:
            raise ValueError('out_channels must be divisible by groups')
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.kernel_size = _pair(kernel_size)
        self.stride = _pair(stride)
        self.dilation = _pair(dilation)
        self.groups = groups
        self.hyper_params = np.prod((out_channels, in_channels // groups) + self.kernel_size)
        self._reset_head()

    def forward(self, x, s, fused_norm=None, fused_act=None):
        if self.kernel_size != (1, 1) or self.stride != (1, 1):
            raise NotImplementedError("HyperPatchNoPadding kernels cover kernel_size 1, stride 1 (every reference "
                                      "model builds it that way, hyperseg_v1_0.py:748-750)")
        weight = self.apply_signal2weights(s)
        scale = shift = None
        if fused_norm is not None:
            scale, shift = ops.fold_bn(fused_norm)
        return ops.patch_conv1x1(x, weight, self.out_channels, self.groups, scale, shift, fused_act)


class HyperPatch(nn.Module, _SignalHeadMixin):
    """Generic patch-wise wrapper around a MetaConv2d with halo padding (reference :501-557)."""

    supports_fused_epilogue = True
    accepts_strided_weights = True
    _pad_head_to_groups = False          # reference :534 builds this head without rounding up to groups

    def __init__(self, module: nn.Module, padding=0, padding_mode='reflect'):
        super().__init__()
        if padding_mode not in _PADDING_MODES:
            raise ValueError(f"padding_mode must be one of {set(_PADDING_MODES)}, but got padding_mode='{padding_mode}'")
        if not isinstance(module, MetaConv2d):
            raise NotImplementedError("hyperseg_b200 HyperPatch wraps MetaConv2d (the only module the reference wraps)")
        self.hyper_module = module
        self.padding = _pair(padding)
        self.padding_mode = padding_mode
        self._padding_repeated_twice = self.padding + self.padding
        self._reset_head()

    @property
    def hyper_params(self):
        return self.hyper_module.hyper_params

    def forward(self, x, s, fused_norm=None, fused_act=None):
        weight = self.apply_signal2weights(s)
        return _run_patch_conv(self.hyper_module, self.padding, self.padding_mode, x, weight, fused_norm, fused_act)


class HyperPatchConv2d(HyperPatch):
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


def make_hyper_patch_conv2d_block(in_nc, out_nc, kernel_size=3, stride=1, padding=None, dilation=1, groups=1,
                                  padding_mode='reflect', norm_layer=nn.BatchNorm2d, act_layer=nn.ReLU(True),
                                  dropout=None):
    """Patch-wise conv [+ norm] [+ activation] [+ dropout] in a MetaSequential (reference :728-760)."""
    assert dropout is None or isinstance(dropout, float)
    padding = kernel_size // 2 if padding is None else padding
    if padding == 0:
        layers = [HyperPatchNoPadding(in_nc, out_nc, kernel_size, stride, dilation, groups)]
    else:
        layers = [HyperPatchConv2d(in_nc, out_nc, kernel_size, stride, padding, dilation, groups, padding_mode)]
    if norm_layer is not None:
        layers.append(norm_layer(out_nc))
    if act_layer is not None:
        layers.append(act_layer)
    if dropout is not None:
        layers.append(nn.Dropout(dropout))
    return MetaSequential(*layers)


_HEAD_OWNERS = (HyperPatchConv2d, HyperPatchNoPadding, HyperPatchInvertedResidual)


def get_hyper_params(model):
    """Depth-first list of the weight counts of all head-owning layers (reference :256-266)."""
    counts = []
    for child in model.children():
        if isinstance(child, _HEAD_OWNERS):
            counts.append(child.hyper_params)
        else:
            counts += get_hyper_params(child)
    return counts


def init_signal2weights(model, signal_features, signal_index=0, weight_groups=1):
    """Create every layer's head (reference :269-278).

    Faithful to the reference's traversal: the running signal offset is advanced only among *sibling*
    head owners; a recursive call gets the offset by value and its advance is not propagated back.  In the
    shipped models each decoder level is a separate container, so every head ends up with signal_index 0."""
    for child in model.children():
        if isinstance(child, _HEAD_OWNERS):
            channels = signal_features.pop(0)
            groups = weight_groups.pop(0) if isinstance(weight_groups, list) else weight_groups
            child.init_signal2weights(channels, signal_index, groups)
            signal_index += channels
        else:
            init_signal2weights(child, signal_features, signal_index, weight_groups)


def divide_feature(in_feature, out_features, min_unit=8):
    """Split ``in_feature`` channels between consumers proportionally to ``out_features`` in multiples of
    ``min_unit`` (reference :763-810).  Equal consumers get equal shares; consumer classes are served in
    decreasing order of total demand, each receiving floor(share) minus one unit per member beyond the one
    unit every consumer is guaranteed; the last class takes what is left."""
    assert in_feature % min_unit == 0, f'in_feature ({in_feature}) must be divisible by min_unit ({min_unit})'
    units = in_feature // min_unit
    order = np.argsort(out_features)
    ranked = np.array(out_features)[order]
    classes = [(size, order[list(members)]) for size, members in groupby(range(len(order)), lambda i: ranked[i])]
    classes.sort(key=lambda c: c[0] * len(c[1]), reverse=True)
    units_per_weight = float(units) / sum(out_features)

    class_units = [len(members) for _, members in classes]      # one unit each, to start with
    spare = units - sum(class_units)
    for ci, (size, members) in enumerate(classes):
        if ci == len(classes) - 1:
            class_units[-1] += spare
            break
        n = len(members)
        share = max(size * n * units_per_weight, n)
        share = share // n * n - n
        share = min(share, spare)
        class_units[ci] += share
        spare -= share
        if spare == 0:
            break

    result = np.zeros(len(out_features), dtype=int)
    for cu, (_, members) in zip(class_units, classes):
        for m in members:
            result[m] = cu // len(members) * min_unit
    return result


def assemble_level_input(coords, skip, p):
    """A decoder level's input: [coordinates | encoder feature | previous level, bilinearly upsampled]
    (reference :235-240).  On the GPU this is one hsb_decoder_input_fwd launch; the stock PyTorch sequence is kept
    for CPU tensors (the decoder itself has no CPU path -- this only serves host-side tests of the plumbing)."""
    if skip.is_cuda and not (torch.is_grad_enabled() and (skip.requires_grad or (p is not None and p.requires_grad))):
        return ops.decoder_input(coords, skip, p)
    if p is not None:
        if p.shape[2:] != skip.shape[2:]:
            p = F.interpolate(p, skip.shape[2:], mode='bilinear', align_corners=False)
        p = torch.cat((skip, p), dim=1)
    else:
        p = skip
    return torch.cat([coords.expand(p.shape[0], -1, -1, -1).to(p.dtype), p], dim=1)


class MultiScaleDecoder(nn.Module):
    """Coarse-to-fine decoder whose blocks are patch-wise dynamic layers (reference :94-253)."""

    def __init__(self, feat_channels, signal_channels, num_classes=3, kernel_sizes=3, level_layers=1,
                 level_channels=None, norm_layer=nn.BatchNorm2d, act_layer=nn.ReLU6(inplace=True), out_kernel_size=1,
                 expand_ratio=1, groups=1, weight_groups=1, with_out_fc=False, dropout=None, coords_res=None):
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
        self.layer_params = []
        self.coords_cache = {}
        self.weight_groups = weight_groups
        feat_channels = feat_channels[::-1]          # coarse -> fine

        carried = 0                                   # channels handed from the previous level
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
            self.add_module(f'level_{level}', MetaSequential(*blocks))

        if with_out_fc:
            tail = [nn.Dropout2d(dropout, True)] if dropout is not None else []
            tail.append(HyperPatchConv2d(carried, num_classes, out_kernel_size, padding=out_kernel_size // 2))
            self.out_fc = MetaSequential(*tail)
        else:
            self.out_fc = None

        self.hyper_params = 0
        self._ranges = [0]
        self.param_groups = []
        for level in range(n_levels):
            hp = getattr(self, f'level_{level}').hyper_params
            self.hyper_params += hp
            self._ranges.append(self.hyper_params)
            self.param_groups.append(hp)
        if with_out_fc:
            self.hyper_params += self.out_fc.hyper_params
            self.param_groups.append(self.out_fc.hyper_params)
        self._ranges.append(self.hyper_params)

        if coords_res is not None:
            for res in coords_res:
                for i in range(n_levels):
                    h, w = res[0] // 2 ** i, res[1] // 2 ** i
                    self.register_buffer(f'coord{h}_{w}', self.cache_image_coordinates(h, w))

        hyper_params = get_hyper_params(self)
        min_unit = max(weight_groups)
        signal_features = divide_feature(signal_channels, hyper_params, min_unit=min_unit)
        head_groups = list(weight_groups) if isinstance(weight_groups, list) else weight_groups
        init_signal2weights(self, list(signal_features), weight_groups=head_groups)
        self.hyper_params = sum(hyper_params)

    @staticmethod
    def _coordinate_grid(h, w, device=None):
        xs = torch.linspace(-1, 1, steps=w, device=device)
        ys = torch.linspace(-1, 1, steps=h, device=device)
        return torch.stack((xs.view(1, w).expand(h, w), ys.view(h, 1).expand(h, w)), dim=0).unsqueeze(0)

    def cache_image_coordinates(self, h, w):
        return self._coordinate_grid(h, w).contiguous()

    def get_image_coordinates(self, b, h, w, device):
        cached = getattr(self, f'coord{h}_{w}', None)
        grid = cached if cached is not None else self._coordinate_grid(h, w, device)
        return grid.expand(b, -1, -1, -1)

    def forward_features(self, x, s):
        """Everything but the final upsample to the frame resolution."""
        p = None
        for level in range(self.levels):
            skip = x[-level - 1]
            coords = self.get_image_coordinates(1, *skip.shape[-2:], skip.device)
            p = assemble_level_input(coords, skip, p)
            p = getattr(self, f'level_{level}')(p, s)
        if self.out_fc is not None:
            p = self.out_fc(p, s)
        return p

    def forward(self, x, s):
        p = self.forward_features(x, s)
        if p.shape[2:] != x[0].shape[2:]:
            p = F.interpolate(p, x[0].shape[2:], mode='bilinear', align_corners=False)
        return p


class HyperGen(nn.Module):
    """Backbone -> weight mapper -> dynamic decoder (reference :12-91)."""

    def __init__(self, backbone, weight_mapper, in_nc=3, num_classes=3, kernel_sizes=3, level_layers=1,
                 level_channels=None, expand_ratio=1, groups=1, weight_groups=1, inference_hflip=False,
                 inference_gather='mean', with_out_fc=False, decoder_groups=1, decoder_dropout=None, coords_res=None):
        super().__init__()
        self.inference_hflip = inference_hflip
        self.inference_gather = inference_gather
        self.backbone = backbone()
        feat_channels = [in_nc] + self.backbone.feat_channels[:-1]
        self.decoder = MultiScaleDecoder(feat_channels, self.backbone.feat_channels[-1], num_classes, kernel_sizes,
                                         level_layers, level_channels, with_out_fc=with_out_fc, out_kernel_size=1,
                                         expand_ratio=expand_ratio, groups=decoder_groups, weight_groups=weight_groups,
                                         dropout=decoder_dropout, coords_res=coords_res)
        self.weight_mapper = weight_mapper(self.backbone.feat_channels[-1], self.decoder.param_groups)

    @property
    def hyper_params(self):
        return self.decoder.hyper_params

    def process_single_tensor(self, x, hflip=False):
        if hflip:
            x = torch.flip(x, [-1])
        features = self.backbone(x)
        signal = self.weight_mapper(features[-1])
        out = self.decoder([x] + features[:-1], signal)
        return torch.flip(out, [-1]) if hflip else out

    def gather_results(self, x, y=None):
        assert x is not None
        if y is None:
            return x
        return (x + y) * 0.5 if self.inference_gather == 'mean' else torch.max(x, y)

    def forward(self, x):
        assert isinstance(x, (list, tuple, torch.Tensor)), 'x must be of type list, tuple, or tensor'
        if isinstance(x, torch.Tensor):
            return self.process_single_tensor(x)
        # image pyramid (+ optional horizontal-flip TTA); the first entry fixes the output resolution
        out_res = x[0].shape[2:]
        out = None
        for img in x:
            pred = self.process_single_tensor(img)
            if self.inference_hflip:
                pred = torch.max(pred, self.process_single_tensor(img, hflip=True))
            if pred.shape[2:] != out_res:
                pred = F.interpolate(pred, out_res, mode='bilinear', align_corners=False)
            out = self.gather_results(pred, out)
        return out


def hyperseg_efficientnet(model_name, pretrained=False, out_feat_scale=0.25, levels=3, weights_path=None, **kwargs):
    """Factory with the reference's signature (reference :813-827)."""
    from .efficientnet import efficientnet
    weight_mapper = partial(WeightMapper, levels=levels)
    backbone = partial(efficientnet, model_name, pretrained=pretrained, out_feat_scale=out_feat_scale, head=None,
                       return_features=True)
    model = HyperGen(backbone, weight_mapper, **kwargs)
    if weights_path is not None:
        checkpoint = torch.load(weights_path, map_location='cpu')
        model.load_state_dict(checkpoint['state_dict'], strict=True)
    return model
