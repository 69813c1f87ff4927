"""GPU BASELINE of the hot path, test/bench infrastructure only (never imported by hyperseg_b200).

The reference runs its decoder through stock ATen: reflect pad, ``unfold`` into halo tiles, weight reshapes and
``F.conv2d(groups = batch * patches)``, BatchNorm and ReLU6 as separate kernels, ``permute`` copies back to NCHW
(hyperseg/models/hyperseg_v1_0.py:328-370 inverted-residual block, :486-498 1x1 block, :315-326 weight head).
This file restates that *operator sequence* with the same torch calls, so that the same B200 that runs libhsb200
can also time what the reference would do there (bench.py ``gpu_reference``), in fp32 and under bf16 autocast --
the honest GPU baseline SURVEY section 8(d) asks for.  It is written from the survey's index-level description
(Appendix A), not copied from the reference sources, and it is checked against the float64 oracle in
tests/test_oracle_golden.py.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def head(s, weight, sig_index, sig_ch, hp, groups):
    """Grouped 1x1 convolution of the signal slice, truncated to the hp weights the layer uses."""
    return F.conv2d(s[:, sig_index:sig_index + sig_ch], weight, None, groups=groups)[:, :hp]


def _rows(w):
    """(B, hp, fh, fw) -> one weight vector per (image, patch): (B*fh*fw, hp)."""
    return w.permute(0, 2, 3, 1).reshape(-1, w.shape[1])


def patch_conv1x1(x, w, out_channels, bn=None, relu=True):
    """Per-patch 1x1 convolution as ONE grouped convolution with a group per (image, patch)."""
    B, C, H, W = x.shape
    fh, fw = w.shape[-2:]
    ph, pw = H // fh, W // fw
    tiles = x.reshape(B, C, fh, ph, fw, pw).permute(0, 2, 4, 1, 3, 5).reshape(1, B * fh * fw * C, ph, pw)
    k = _rows(w).reshape(B * fh * fw * out_channels, C, 1, 1)
    y = F.conv2d(tiles, k, None, groups=B * fh * fw)
    y = y.reshape(B, fh, fw, out_channels, ph, pw).permute(0, 3, 1, 4, 2, 5).reshape(B, out_channels, H, W)
    if bn is not None:
        y = bn(y)
    return F.relu(y) if relu else y


def patch_ir(x, w, hidden, out_channels, bn1, bn2, bn3):
    """Inverted-residual MetaBlock on halo tiles: pad, unfold, three grouped convolutions with BatchNorm / ReLU6
    between them, and the re-tiling copy."""
    B, C, H, W = x.shape
    fh, fw = w.shape[-2:]
    ph, pw = H // fh, W // fw
    th, tw = ph + 2, pw + 2
    n = B * fh * fw
    xp = F.pad(x, (1, 1, 1, 1), mode="reflect")
    tiles = xp.unfold(2, th, ph).unfold(3, tw, pw)                       # (B, C, fh, fw, th, tw)
    tiles = tiles.permute(0, 2, 3, 1, 4, 5).reshape(1, n * C, th, tw)
    rows = _rows(w)
    r1, r2 = C * hidden, C * hidden + 9 * hidden
    h = F.conv2d(tiles, rows[:, :r1].reshape(n * hidden, C, 1, 1), None, groups=n)
    h = F.relu6(bn1(h.reshape(n, hidden, th, tw))).reshape(1, n * hidden, th, tw)
    d = F.conv2d(h, rows[:, r1:r2].reshape(n * hidden, 1, 3, 3), None, groups=n * hidden)
    d = F.relu6(bn2(d.reshape(n, hidden, ph, pw))).reshape(1, n * hidden, ph, pw)
    o = F.conv2d(d, rows[:, r2:].reshape(n * out_channels, hidden, 1, 1), None, groups=n)
    o = bn3(o.reshape(n, out_channels, ph, pw))
    return o.reshape(B, fh, fw, out_channels, ph, pw).permute(0, 3, 1, 4, 2, 5).reshape(B, out_channels, H, W)


def make_bn(channels, scale, shift, device):
    """Eval-mode BatchNorm2d whose folded form is y = scale * x + shift."""
    bn = torch.nn.BatchNorm2d(channels).to(device).eval()
    with torch.no_grad():
        bn.running_mean.zero_()
        bn.running_var.fill_(1.0 - bn.eps)
        bn.weight.copy_(scale)
        bn.bias.copy_(shift)
    return bn
