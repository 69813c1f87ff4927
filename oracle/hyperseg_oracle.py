"""CPU oracle for HyperSeg's decoder hot path -- TEST INFRASTRUCTURE, not product code.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
module; nothing under hyperseg_b200/ does, and the product has no CPU path.

This is a restatement, written from the index-level semantics of the reference,
of what the reference computes with pad/unfold/permute/F.conv2d(groups=B*P)/BatchNorm/fold.  Each function
names the reference lines it follows.  It uses explicit tile gathers and einsum contractions -- not
F.conv2d / F.unfold / F.fold / F.pad -- so that it is an independent statement of the arithmetic, and it
accumulates in float64 by default.

Parity pinning: the reference holds no golden vectors or known-answer tests, so this
oracle is pinned against outputs of the reference itself, generated in the build container by
tests/golden/make_golden.py (which imports /root/reference) and committed under tests/golden/*.npz;
tests/test_oracle_golden.py checks every function here against them.

The function signatures equal those of hyperseg_b200.ops so a test can swap one for the other.
"""
from __future__ import annotations

import torch

DTYPE = torch.float64     # accumulation / internal dtype; bench's CPU leg switches this to float32

_ACT = {
    "none": lambda t: t, None: (lambda t: t),
    "relu": lambda t: t.clamp_min(0),
    "relu6": lambda t: t.clamp(0, 6),
}


# ---- padding by index arithmetic (reference: F.pad(..., mode), hyperseg_v1_0.py:337, meta_patch.py:50) ----
def _in(t):
    """Inputs are detached unless a gradient is being taken through the oracle (training-path tests)."""
    if torch.is_grad_enabled() and t.requires_grad:
        return t.to(DTYPE)
    return t.detach().to(DTYPE)


def _source_index(n, pad, mode):
    """For coordinates -pad .. n+pad-1: source index in [0, n) and validity mask."""
    i = torch.arange(-pad, n + pad)
    valid = torch.ones_like(i, dtype=torch.bool)
    if mode == "reflect":            # mirror without repeating the border sample
        if n > 1:
            period = 2 * (n - 1)
            m = torch.remainder(i, period)
            i = torch.where(m < n, m, period - m)
        else:
            i = torch.zeros_like(i)
    elif mode == "replicate":
        i = i.clamp(0, n - 1)
    elif mode == "circular":
        i = torch.remainder(i, n)
    elif mode == "zeros":
        valid = (i >= 0) & (i < n)
        i = i.clamp(0, n - 1)
    else:
        raise ValueError(mode)
    return i, valid


def pad_map(x, pad_h, pad_w, mode):
    iy, vy = _source_index(x.shape[-2], pad_h, mode)
    ix, vx = _source_index(x.shape[-1], pad_w, mode)
    out = x[..., iy[:, None], ix[None, :]]
    if mode == "zeros":
        out = out * (vy[:, None] & vx[None, :]).to(out.dtype)
    return out


def gather_tiles(xp, fh, fw, ph, pw, th, tw):
    """tiles[b,c,i,j,r,q] = xp[b,c,i*ph+r,j*pw+q]  (reference: unfold(1,kh,ph).unfold(2,kw,pw), hyperseg_v1_0.py:342;
    F.unfold((kh,kw), stride=(ph,pw)), meta_patch.py:51)."""
    rows = torch.arange(fh)[:, None] * ph + torch.arange(th)[None, :]       # (fh, th)
    cols = torch.arange(fw)[:, None] * pw + torch.arange(tw)[None, :]       # (fw, tw)
    return xp[:, :, rows[:, None, :, None], cols[None, :, None, :]]          # (B, C, fh, fw, th, tw)


def scatter_tiles(t):
    """(B, C, fh, fw, ph, pw) -> (B, C, fh*ph, fw*pw)  (reference: view/permute/reshape hyperseg_v1_0.py:368, :496;
    F.fold with non-overlapping tiles, meta_patch.py:55)."""
    B, C, fh, fw, ph, pw = t.shape
    return t.permute(0, 1, 2, 4, 3, 5).reshape(B, C, fh * ph, fw * pw)


def _patch_rows(w):
    """(B, hp, fh, fw) -> (B, fh, fw, hp)  (reference: weight.permute(0,2,3,1), hyperseg_v1_0.py:345-347, :492, :549)."""
    return w.permute(0, 2, 3, 1)


def _epilogue(y, scale, shift, act):
    if scale is not None:
        y = y * scale.to(y.dtype).view(1, -1, 1, 1) + shift.to(y.dtype).view(1, -1, 1, 1)
    return _ACT[act](y)


def fold_bn(bn):
    """Eval-mode BatchNorm2d as scale/shift: (t - mean)/sqrt(var + eps)*gamma + beta."""
    var = bn.running_var.to(DTYPE)
    scale = 1.0 / torch.sqrt(var + bn.eps)
    if bn.weight is not None:
        scale = scale * bn.weight.detach().to(DTYPE)
    shift = -bn.running_mean.to(DTYPE) * scale
    if bn.bias is not None:
        shift = shift + bn.bias.detach().to(DTYPE)
    return scale, shift


# ---- a1: HyperPatchNoPadding.forward, hyperseg_v1_0.py:486-498 ------------------------------------------
def patch_conv1x1(x, w, out_channels, groups=1, scale=None, shift=None, act="none"):
    out_dtype = x.dtype
    x, w = _in(x), _in(w)
    B, Cin, H, W = x.shape
    fh, fw = w.shape[-2:]
    ph, pw = H // fh, W // fw
    cig, cog = Cin // groups, out_channels // groups
    Wm = _patch_rows(w).reshape(B, fh, fw, groups, cog, cig)            # Wm[o,c] = w[o*cig + c]
    xt = x.reshape(B, groups, cig, fh, ph, fw, pw)
    y = torch.einsum('bijgoc,bgciujv->bgoiujv', Wm, xt).reshape(B, out_channels, H, W)
    return _epilogue(y, scale, shift, act).to(out_dtype)


# ---- a2: HyperPatchInvertedResidual.conv / forward, hyperseg_v1_0.py:328-376 ----------------------------
def patch_ir(x, w, hidden, out_channels, bn1, bn2, bn3, residual=False):
    out_dtype = x.dtype
    x, w = _in(x), _in(w)
    B, Cin, H, W = x.shape
    fh, fw = w.shape[-2:]
    ph, pw = H // fh, W // fw
    r1 = Cin * hidden
    r2 = r1 + 9 * hidden
    rows = _patch_rows(w)                                               # (B, fh, fw, hp)
    W1 = rows[..., :r1].reshape(B, fh, fw, hidden, Cin)                 # :350
    W2 = rows[..., r1:r2].reshape(B, fh, fw, hidden, 3, 3)              # :357-358
    W3 = rows[..., r2:].reshape(B, fh, fw, out_channels, hidden)        # :364
    tiles = gather_tiles(pad_map(x, 1, 1, "reflect"), fh, fw, ph, pw, ph + 2, pw + 2)   # :337, :342
    bc = lambda v: v.to(DTYPE).view(1, -1, 1, 1, 1, 1)
    h = torch.einsum('bijoc,bcijrq->boijrq', W1, tiles)                 # :351
    h = (h * bc(bn1[0]) + bc(bn1[1])).clamp(0, 6)                       # :352-353
    d = torch.zeros(B, hidden, fh, fw, ph, pw, dtype=DTYPE)
    for ky in range(3):                                                 # :359 depthwise, valid
        for kx in range(3):
            d = d + W2[..., ky, kx].permute(0, 3, 1, 2)[..., None, None] * h[..., ky:ky + ph, kx:kx + pw]
    d = (d * bc(bn2[0]) + bc(bn2[1])).clamp(0, 6)                       # :360-361
    o = torch.einsum('bijoc,bcijuv->boijuv', W3, d)                     # :365
    o = o * bc(bn3[0]) + bc(bn3[1])                                     # :366
    y = scatter_tiles(o)                                                # :368
    if residual:
        y = y + x                                                       # :373-374
    return y.to(out_dtype)


# ---- a3/a4/a8: signal2weights heads, hyperseg_v1_0.py:315-326; unify :287-309; v0_1 :336-362 --------------
def signal2weights(s, ws, sig_index, sig_ch, hp, groups):
    out_dtype = s.dtype
    s, ws = _in(s), _in(ws)
    B, _, fh, fw = s.shape
    out_ch = ws.shape[0]
    spg, opg = sig_ch // groups, out_ch // groups
    sl = s[:, sig_index:sig_index + sig_ch].reshape(B, groups, spg, fh, fw)
    Wg = ws.reshape(groups, opg, spg)
    full = torch.einsum('gok,bgkij->bgoij', Wg, sl).reshape(B, out_ch, fh, fw)
    return full[:, :hp].to(out_dtype)


# ---- a6/a7: MetaPatch.forward + MetaConv2d.forward, meta_patch.py:35-57, meta_conv.py:163-186 -------------
def patch_conv(x, w, out_channels, kernel_size, padding, dilation=(1, 1), groups=1, padding_mode="reflect",
               scale=None, shift=None, act="none"):
    out_dtype = x.dtype
    x, w = _in(x), _in(w)
    B, Cin, H, W = x.shape
    fh, fw = w.shape[-2:]
    ph, pw = H // fh, W // fw
    kh, kw = kernel_size
    cig, cog = Cin // groups, out_channels // groups
    th, tw = ph + 2 * padding[0], pw + 2 * padding[1]
    oh, ow = th - dilation[0] * (kh - 1), tw - dilation[1] * (kw - 1)
    assert (oh, ow) == (ph, pw), "only size-preserving geometry is used by the reference"
    tiles = gather_tiles(pad_map(x, padding[0], padding[1], padding_mode), fh, fw, ph, pw, th, tw)
    tiles = tiles.reshape(B, groups, cig, fh, fw, th, tw)
    Wm = _patch_rows(w).reshape(B, fh, fw, groups, cog, cig, kh, kw)
    y = torch.zeros(B, groups, cog, fh, fw, ph, pw, dtype=DTYPE)
    for ky in range(kh):
        for kx in range(kw):
            win = tiles[..., ky * dilation[0]:ky * dilation[0] + ph, kx * dilation[1]:kx * dilation[1] + pw]
            y = y + torch.einsum('bijgoc,bgcijuv->bgoijuv', Wm[..., ky, kx], win)
    y = scatter_tiles(y.reshape(B, out_channels, fh, fw, ph, pw))
    return _epilogue(y, scale, shift, act).to(out_dtype)


def meta_conv2d(x, w, out_channels, kernel_size, padding=(0, 0), dilation=(1, 1), groups=1, padding_mode="zeros"):
    """Per-sample dynamic convolution (meta_conv.py:163-186), stride 1."""
    out_dtype = x.dtype
    x, w = _in(x), _in(w)
    N, Cin, H, W = x.shape
    kh, kw = kernel_size
    cig, cog = Cin // groups, out_channels // groups
    xp = pad_map(x, padding[0], padding[1], padding_mode).reshape(N, groups, cig, H + 2 * padding[0], W + 2 * padding[1])
    Ho = H + 2 * padding[0] - dilation[0] * (kh - 1)
    Wo = W + 2 * padding[1] - dilation[1] * (kw - 1)
    Wm = w.reshape(N, groups, cog, cig, kh, kw)
    y = torch.zeros(N, groups, cog, Ho, Wo, dtype=DTYPE)
    for ky in range(kh):
        for kx in range(kw):
            win = xp[..., ky * dilation[0]:ky * dilation[0] + Ho, kx * dilation[1]:kx * dilation[1] + Wo]
            y = y + torch.einsum('ngoc,ngcuv->ngouv', Wm[..., ky, kx], win)
    return y.reshape(N, out_channels, Ho, Wo).to(out_dtype)


def weights_to_patch_major(w):
    return w


FUNCTIONS = ("patch_conv1x1", "patch_ir", "signal2weights", "patch_conv", "meta_conv2d", "fold_bn",
             "weights_to_patch_major")


class use_oracle_ops:
    """Context manager for tests / the CPU-baseline leg: route hyperseg_b200.ops.* to this module so the
    nn.Module mirror can be executed on CPU.  The product never does this."""

    def __init__(self, dtype=torch.float64):
        self.dtype = dtype

    def __enter__(self):
        global DTYPE
        import hyperseg_b200.ops as ops
        self._ops = ops
        self._saved = {n: getattr(ops, n) for n in FUNCTIONS}
        self._saved_dtype = DTYPE
        DTYPE = self.dtype
        g = globals()
        for n in FUNCTIONS:
            setattr(ops, n, g[n])
        return self

    def __exit__(self, *exc):
        global DTYPE
        for n, f in self._saved.items():
            setattr(self._ops, n, f)
        DTYPE = self._saved_dtype
        return False
