"""Tensor-level wrappers over the C ABI (include/hsb200.h).

PyTorch is used here only for device memory and the current stream: every function hands raw device
pointers to libhsb200.so and returns a tensor it allocated for the output.  Inputs must live on a CUDA
device -- there is no CPU implementation to fall back to, and none is attempted.

Per-patch weight tensors are exchanged as logical ``(B, hp, fh, fw)`` tensors (the reference's shape).
Their memory may be either contiguous NCHW (the reference layout) or "patch-major", i.e. what
PyTorch calls channels_last -- ``(B, fh, fw, row)`` storage with ``row >= hp`` -- which is what
:func:`signal2weights` produces and what the kernels stream without a layout change.
"""
from __future__ import annotations

import weakref

import torch

from . import _lib
from ._lib import ACT_NONE, ACT_RELU, ACT_RELU6, ACT_SILU, HSB_BF16, HSB_F32, PAD_MODES, W_NCHW, W_PATCH_MAJOR

import os

_NO_TC_HEADS = os.environ.get("HSB_DISABLE_TC_HEADS", "0") == "1"
_LAUNCHES = 0      # kernels launched through the C ABI by this process (each entry point launches exactly one)


def launch_count() -> int:
    return _LAUNCHES


def _call(fn_name, *args):
    global _LAUNCHES
    _lib.check(getattr(_lib.load(), fn_name)(*args), fn_name)
    _LAUNCHES += 1


ACTS = {"none": ACT_NONE, None: ACT_NONE, "relu": ACT_RELU, "relu6": ACT_RELU6, "silu": ACT_SILU}
_DTYPES = {torch.float32: HSB_F32, torch.bfloat16: HSB_BF16}


def _compute_dtype(x: torch.Tensor) -> torch.dtype:
    if x.is_cuda and torch.is_autocast_enabled("cuda"):
        dt = torch.get_autocast_dtype("cuda")
    else:
        dt = x.dtype
    if dt not in _DTYPES:
        raise TypeError(f"hyperseg_b200 kernels support float32 and bfloat16, got {dt}")
    return dt


def _require_cuda(*tensors: torch.Tensor) -> None:
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError(
                "hyperseg_b200 operators run only on CUDA tensors (sm_100a kernels, no CPU fallback); "
                f"got a tensor on {t.device}")


def _needs_grad(*tensors) -> bool:
    return torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in tensors)


def _require_inference(*tensors: torch.Tensor) -> None:
    """For the operators that have no backward kernel (the fused inverted-residual block, the glue kernels)."""
    if _needs_grad(*tensors):
        raise NotImplementedError(
            "this hyperseg_b200 operator is forward-only: gradients are implemented for the patch-wise convolutions "
            "(MetaPatchConv2d / HyperPatchNoPadding / HyperPatchConv2d) and the weight heads; the module "
            "HyperPatchInvertedResidual differentiates through its stage-wise path, the fused kernel (ops.patch_ir) "
            "itself must run under torch.no_grad() / on inputs that do not require grad")


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _fptr(t):
    return None if t is None else t.data_ptr()


def _affine(scale, shift, channels, device):
    if scale is None and shift is None:
        return None, None
    if scale is None or shift is None:
        raise ValueError("scale and shift must be given together")
    scale = scale.to(device=device, dtype=torch.float32).contiguous()
    shift = shift.to(device=device, dtype=torch.float32).contiguous()
    if scale.numel() != channels or shift.numel() != channels:
        raise ValueError(f"epilogue scale/shift must have {channels} entries")
    return scale, shift


def weight_layout(w: torch.Tensor):
    """Classify a logical (B, hp, fh, fw) weight tensor -> (tensor, layout code, row stride)."""
    if w.dim() != 4:
        raise ValueError(f"per-patch weights must be 4-D (B, hp, fh, fw), got {tuple(w.shape)}")
    B, hp, fh, fw = w.shape
    st = w.stride()
    if fw > 1:
        row = st[3]
    elif fh > 1:
        row = st[2]
    elif B > 1:
        row = st[0]
    else:
        row = hp if (hp == 1 or st[1] == 1) else -1
    ok = row >= hp
    ok = ok and (hp == 1 or st[1] == 1)
    ok = ok and (fw == 1 or st[3] == row)
    ok = ok and (fh == 1 or st[2] == fw * row)
    ok = ok and (B == 1 or st[0] == fh * fw * row)
    if ok:
        return w, W_PATCH_MAJOR, int(row)
    if not w.is_contiguous():
        w = w.contiguous()
    return w, W_NCHW, 0


def fold_bn(bn: torch.nn.BatchNorm2d):
    """Eval-mode BatchNorm2d as y = scale * x + shift (fp32).

    An inference engine may pin the result on the module (``_hsb_folded``) so that it is computed once, in fp32,
    before the module's buffers are cast to a 16-bit dtype."""
    cached = getattr(bn, "_hsb_folded", None)
    if cached is not None:
        return cached
    if bn.training or bn.running_mean is None:
        raise NotImplementedError("only eval-mode BatchNorm with running statistics can be fused")
    var = bn.running_var.float()
    scale = torch.rsqrt(var + bn.eps)
    if bn.weight is not None:
        scale = scale * bn.weight.float()
    shift = -bn.running_mean.float() * scale
    if bn.bias is not None:
        shift = shift + bn.bias.float()
    return scale.contiguous(), shift.contiguous()


def _prep_xw(x, w):
    _require_cuda(x, w)
    dt = _compute_dtype(x)
    x = x.to(dt).contiguous()
    if w.dtype != dt:
        w = w.to(dt)
    w, layout, row = weight_layout(w)
    if x.shape[0] != w.shape[0]:
        raise ValueError(f"batch mismatch between x {tuple(x.shape)} and weights {tuple(w.shape)}")
    return x, w, layout, row, dt


def _unfused_epilogue(y, scale, shift, act):
    if scale is not None:
        y = y * scale.to(y.dtype).view(1, -1, 1, 1) + shift.to(y.dtype).view(1, -1, 1, 1)
    if act == "relu":
        y = torch.relu(y)
    elif act == "relu6":
        y = torch.nn.functional.relu6(y)
    return y


def patch_conv1x1(x, w, out_channels, groups=1, scale=None, shift=None, act="none"):
    """Patch-wise 1x1 convolution (+ fused per-channel affine and activation)."""
    if _needs_grad(x, w, scale, shift):
        y = _PatchConvFn.apply(x, w, out_channels, (1, 1), (0, 0), (1, 1), groups, "zeros")
        return _unfused_epilogue(y, scale, shift, act)
    return _patch_conv1x1_fwd(x, w, out_channels, groups, scale, shift, act)


def _patch_conv1x1_fwd(x, w, out_channels, groups=1, scale=None, shift=None, act="none"):
    x, w, layout, row, dt = _prep_xw(x, w)
    B, Cin, H, W = x.shape
    hp = out_channels * (Cin // groups)
    if w.shape[1] != hp:
        raise ValueError(f"expected {hp} weights per patch, got {w.shape[1]}")
    fh, fw = w.shape[-2:]
    scale, shift = _affine(scale, shift, out_channels, x.device)
    y = torch.empty((B, out_channels, H, W), dtype=dt, device=x.device)
    _call("hsb_patch_conv1x1_fwd", x.data_ptr(), w.data_ptr(), y.data_ptr(), _fptr(scale), _fptr(shift), ACTS[act],
        B, Cin, out_channels, H, W, fh, fw, groups, _DTYPES[dt], layout, row, _stream())
    return y


def patch_ir(x, w, hidden, out_channels, bn1, bn2, bn3, residual=False):
    """Fused patch-wise inverted residual block; bn* are (scale, shift) pairs of the folded BatchNorms."""
    _require_inference(x, w)
    x, w, layout, row, dt = _prep_xw(x, w)
    B, Cin, H, W = x.shape
    hp = Cin * hidden + 9 * hidden + hidden * out_channels
    if w.shape[1] != hp:
        raise ValueError(f"expected {hp} weights per patch, got {w.shape[1]}")
    fh, fw = w.shape[-2:]
    s1, b1 = _affine(bn1[0], bn1[1], hidden, x.device)
    s2, b2 = _affine(bn2[0], bn2[1], hidden, x.device)
    s3, b3 = _affine(bn3[0], bn3[1], out_channels, x.device)
    y = torch.empty((B, out_channels, H, W), dtype=dt, device=x.device)
    _call("hsb_patch_ir_fwd", x.data_ptr(), w.data_ptr(), y.data_ptr(), s1.data_ptr(), b1.data_ptr(), s2.data_ptr(), b2.data_ptr(),
        s3.data_ptr(), b3.data_ptr(), B, Cin, hidden, out_channels, H, W, fh, fw, int(bool(residual)),
        _DTYPES[dt], layout, row, _stream())
    return y


def ir_arranged_supported(cin, hidden, out_channels, patch_size) -> bool:
    """True when the restage-free tensor-core MetaBlock kernel has an instantiation for this shape."""
    return bool(_lib.load().hsb_patch_ir_arranged_supported(int(cin), int(hidden), int(out_channels), int(patch_size)))


def ir_arranged_row(cin, hidden, out_channels) -> int:
    """bf16 elements of one arranged weight row (csrc/ir_arranged.cuh)."""
    return int(_lib.load().hsb_ir_arranged_row_elems(int(cin), int(hidden), int(out_channels)))


def ir_arrange_weights(w, cin, hidden, out_channels, s1, s2, s3):
    """Reference-order per-patch weights (B, hp, fh, fw) -> arranged rows (B, fh, fw, row) bf16 with the three
    BatchNorm scales folded in: the operand order of :func:`patch_ir_arranged`."""
    _require_cuda(w)
    if w.dtype not in _DTYPES:
        w = w.float()
    w, layout, row = weight_layout(w)
    B, hp, fh, fw = w.shape
    if hp != cin * hidden + 9 * hidden + hidden * out_channels:
        raise ValueError(f"expected {cin * hidden + 9 * hidden + hidden * out_channels} weights per patch, got {hp}")
    s1, _ = _affine(s1, s1, hidden, w.device)
    s2, _ = _affine(s2, s2, hidden, w.device)
    s3, _ = _affine(s3, s3, out_channels, w.device)
    out = torch.empty((B, fh, fw, ir_arranged_row(cin, hidden, out_channels)), dtype=torch.bfloat16, device=w.device)
    _call("hsb_ir_arrange_weights", w.data_ptr(), out.data_ptr(), s1.data_ptr(), s2.data_ptr(), s3.data_ptr(),
          B, cin, hidden, out_channels, fh, fw, _DTYPES[w.dtype], layout, row, _stream())
    return out


def patch_ir_arranged(x, w_arranged, hidden, out_channels, b1, b2, b3):
    """Fused MetaBlock on arranged weight rows (B, fh, fw, row): bf16, tcgen05 path only -- raises for shapes without an
    instantiation instead of switching kernels.  b1/b2/b3 are the BatchNorm shifts (the scales live in the rows)."""
    _require_cuda(x, w_arranged)
    _require_inference(x, w_arranged)
    if x.dtype != torch.bfloat16 or w_arranged.dtype != torch.bfloat16:
        raise TypeError("patch_ir_arranged is a bf16 kernel")
    x = x.contiguous()
    B, Cin, H, W = x.shape
    if w_arranged.dim() != 4 or w_arranged.shape[0] != B or w_arranged.stride(3) != 1:
        raise ValueError("arranged weights must be (B, fh, fw, row) with contiguous rows")
    fh, fw, row = w_arranged.shape[1:]
    rs = w_arranged.stride(2) if fw > 1 else (w_arranged.stride(1) if fh > 1 else w_arranged.stride(0))
    if (fw > 1 and fh > 1 and w_arranged.stride(1) != fw * rs) or (B > 1 and w_arranged.stride(0) != fh * fw * rs):
        raise ValueError("arranged weight rows must be uniformly strided")
    if row < ir_arranged_row(Cin, hidden, out_channels):
        raise ValueError("arranged rows are too short for this block")
    _, b1 = _affine(b1, b1, hidden, x.device)
    _, b2 = _affine(b2, b2, hidden, x.device)
    _, b3 = _affine(b3, b3, out_channels, x.device)
    y = torch.empty((B, out_channels, H, W), dtype=torch.bfloat16, device=x.device)
    _call("hsb_patch_ir_arranged_fwd", x.data_ptr(), w_arranged.data_ptr(), y.data_ptr(), b1.data_ptr(), b2.data_ptr(),
          b3.data_ptr(), B, Cin, hidden, out_channels, H, W, fh, fw, rs, _stream())
    return y


class ArrangedHead:
    """Static head weights of one inverted-residual block, packed in the block's operand order with its BatchNorm scales
    folded in (hsb_head_pack_arranged).  Owned by whoever built it (the nn.Module / the engine): the buffers stay alive as
    long as this object does, which is what a captured CUDA graph needs."""

    def __init__(self, ws, sig_index, sig_ch, groups, hp_offset, cin, hidden, cout, s1, s2, s3):
        import ctypes
        _require_cuda(ws)
        ws2d = ws.detach().reshape(ws.shape[0], -1)
        if ws2d.dtype not in _DTYPES:
            ws2d = ws2d.float()
        ws2d = ws2d.contiguous()
        elems, items, kmax = ctypes.c_int64(0), ctypes.c_int(0), ctypes.c_int(0)
        _lib.check(_lib.load().hsb_head_arranged_plan(sig_index, sig_ch, ws2d.shape[0], groups, hp_offset, cin, hidden, cout,
                                                      ctypes.byref(elems), ctypes.byref(items), ctypes.byref(kmax)), "hsb_head_arranged_plan")
        self.sig_index, self.sig_ch, self.geom = int(sig_index), int(sig_ch), (int(cin), int(hidden), int(cout))
        self.items, self.kpad_max, self.row = items.value, kmax.value, ir_arranged_row(cin, hidden, cout)
        self.packed = torch.empty(elems.value, dtype=torch.bfloat16, device=ws2d.device)
        self.table = torch.empty((self.items, 4), dtype=torch.int32, device=ws2d.device)
        s1, _ = _affine(s1, s1, hidden, ws2d.device)
        s2, _ = _affine(s2, s2, hidden, ws2d.device)
        s3, _ = _affine(s3, s3, cout, ws2d.device)
        _call("hsb_head_pack_arranged", ws2d.data_ptr(), self.packed.data_ptr(), self.table.data_ptr(), s1.data_ptr(), s2.data_ptr(),
              s3.data_ptr(), self.sig_index, self.sig_ch, ws2d.shape[0], groups, hp_offset, cin, hidden, cout, _DTYPES[ws2d.dtype], _stream())


def head_tc_ok(s, sig_ch_per_group=0) -> bool:
    """True when the signal map meets the tensor-core heads' requirements (bf16 compute, 8-aligned positions)."""
    B, C, fh, fw = s.shape
    return s.is_cuda and _compute_dtype(s) == torch.bfloat16 and (fh * fw) % 8 == 0 and not _NO_TC_HEADS


def signal2weights_arranged(s, head: ArrangedHead):
    """signal (B, C, fh, fw) -> arranged rows (B, fh, fw, row) of the block ``head`` was packed for."""
    _require_cuda(s)
    _require_inference(s)
    if s.dtype != torch.bfloat16:
        s = s.to(torch.bfloat16)
    B, C, fh, fw = s.shape
    if head.sig_index + head.sig_ch > C:
        raise ValueError(f"signal slice [{head.sig_index}, {head.sig_index + head.sig_ch}) exceeds {C} signal channels")
    st = s.stride()
    sp = st[3] if fw > 1 else (st[2] if fh > 1 else 1)
    if sp != 1 or (fh > 1 and fw > 1 and st[2] != fw) or st[0] % 8 or st[1] % 8 or s.data_ptr() % 16:
        s = s.contiguous()
        st = s.stride()
    out = torch.empty((B, fh, fw, head.row), dtype=torch.bfloat16, device=s.device)
    _call("hsb_signal2weights_arranged_fwd", s.data_ptr(), head.packed.data_ptr(), head.table.data_ptr(), out.data_ptr(), B,
          head.sig_index, head.sig_ch, head.items, head.kpad_max, head.row, fh, fw, st[0], st[1], head.row, _stream())
    return out


# (data_ptr, version, shape, groups, dtype) -> (weak reference to the weight tensor that was packed, packed bf16 operand).
# The weak reference ties an entry to one live tensor object: a different tensor that later lands on the same address
# (same shape, version 0) must not be served the old operand.
_PACKED_HEADS = {}


def _cache_lookup(cache, key, owner):
    hit = cache.get(key)
    if hit is not None and hit[0]() is owner:
        return hit[1]
    return None


def _cache_store(cache, key, owner, value, limit=64):
    """Entries whose owner is alive are never evicted: a captured CUDA graph may have the packed buffer's address baked
    in.  Only entries of dead owners (which can never hit again) are pruned once the table grows."""
    if len(cache) > limit:
        for k in [k for k, (ref, _) in cache.items() if ref() is None]:
            del cache[k]
    cache[key] = (weakref.ref(owner), value)


def _packed_head(ws2d, sig_ch, groups, owner):
    key = (ws2d.data_ptr(), ws2d._version, tuple(ws2d.shape), groups, ws2d.dtype)
    packed = _cache_lookup(_PACKED_HEADS, key, owner)
    if packed is not None:
        return packed
    n = _lib.load().hsb_head_packed_elems(sig_ch, ws2d.shape[0], groups)
    if n <= 0:
        raise ValueError("bad head dimensions")
    packed = torch.empty(n, dtype=torch.bfloat16, device=ws2d.device)
    _call("hsb_head_pack", ws2d.data_ptr(), packed.data_ptr(), None, sig_ch, ws2d.shape[0], groups,
          _DTYPES[ws2d.dtype], _stream())
    _cache_store(_PACKED_HEADS, key, owner, packed)
    return packed


def signal2weights(s, ws, sig_index, sig_ch, hp, groups):
    """Grouped 1x1 head: signal (B, C, fh, fw) -> logical (B, hp, fh, fw) weights in patch-major storage.

    ``ws`` is the nn.Conv2d weight (out_ch, sig_ch/groups, 1, 1); only the first ``hp`` output channels
    are produced (the reference computes all ``out_ch`` and slices)."""
    if _needs_grad(s, ws):
        return _HeadFn.apply(s, ws, sig_index, sig_ch, hp, groups)
    return _signal2weights_fwd(s, ws, sig_index, sig_ch, hp, groups)


def _signal2weights_fwd(s, ws, sig_index, sig_ch, hp, groups):
    _require_cuda(s, ws)
    dt = _compute_dtype(s)
    if s.dtype != dt:
        s = s.to(dt)
    ws_owner = ws                                       # the caller's tensor object (normally the nn.Parameter)
    ws_src = ws.detach().reshape(ws.shape[0], -1)       # packed once per (tensor, version) on the tensor-core path
    ws = ws_src.to(dt).contiguous()
    B, C, fh, fw = s.shape
    out_ch = ws.shape[0]
    if ws.shape[1] * groups != sig_ch:
        raise ValueError("signal2weights weight does not match sig_ch / groups")
    if sig_index + sig_ch > C:
        raise ValueError(f"signal slice [{sig_index}, {sig_index + sig_ch}) exceeds {C} signal channels")
    st = s.stride()
    sp = st[3] if fw > 1 else (st[2] if fh > 1 else 1)
    if (fh > 1 and fw > 1 and st[2] != fw * st[3]) or (dt == torch.bfloat16 and sp != 1 and not _NO_TC_HEADS):
        # the tensor-core head wants position-contiguous (NCHW) signal rows; the signal map is small (1/32 resolution)
        s = s.contiguous()
        st = s.stride()
        sp = st[3] if fw > 1 else (st[2] if fh > 1 else 1)
    row = (hp + 7) // 8 * 8
    buf = torch.empty((B, fh, fw, row), dtype=dt, device=s.device)
    if dt == torch.bfloat16 and (fh * fw) % 8 == 0 and sp == 1 and st[0] % 8 == 0 and st[1] % 8 == 0 \
            and s.data_ptr() % 16 == 0 and (sig_ch // groups) <= 128 and not _NO_TC_HEADS:
        # tensor-core path: static weights packed once into the UMMA operand layout
        packed = _packed_head(ws_src if (ws_src.dtype in _DTYPES and ws_src.is_contiguous()) else ws, sig_ch, groups, ws_owner)
        _call("hsb_signal2weights_packed_fwd", s.data_ptr(), packed.data_ptr(), buf.data_ptr(), B, sig_index, sig_ch,
              out_ch, hp, groups, fh, fw, st[0], st[1], row, _stream())
        return buf[..., :hp].permute(0, 3, 1, 2)
    _call("hsb_signal2weights_fwd", s.data_ptr(), ws.data_ptr(), buf.data_ptr(), B, sig_index, sig_ch, out_ch, hp, groups, fh, fw,
        st[0], st[1], sp, _DTYPES[dt], W_PATCH_MAJOR, row, _stream())
    return buf[..., :hp].permute(0, 3, 1, 2)


def patch_conv(x, w, out_channels, kernel_size, padding, dilation=(1, 1), groups=1, padding_mode="reflect",
               scale=None, shift=None, act="none"):
    """General patch-wise convolution (MetaPatchConv2d / HyperPatchConv2d semantics)."""
    if _needs_grad(x, w, scale, shift):
        y = _PatchConvFn.apply(x, w, out_channels, tuple(kernel_size), tuple(padding), tuple(dilation), groups, padding_mode)
        return _unfused_epilogue(y, scale, shift, act)
    return _patch_conv_fwd(x, w, out_channels, kernel_size, padding, dilation, groups, padding_mode, scale, shift, act)


def _patch_conv_fwd(x, w, out_channels, kernel_size, padding, dilation=(1, 1), groups=1, padding_mode="reflect",
                    scale=None, shift=None, act="none"):
    x, w, layout, row, dt = _prep_xw(x, w)
    B, Cin, H, W = x.shape
    kh, kw = kernel_size
    hp = out_channels * (Cin // groups) * kh * kw
    if w.shape[1] != hp:
        raise ValueError(f"expected {hp} weights per patch, got {w.shape[1]}")
    fh, fw = w.shape[-2:]
    scale, shift = _affine(scale, shift, out_channels, x.device)
    y = torch.empty((B, out_channels, H, W), dtype=dt, device=x.device)
    _call("hsb_patch_conv_fwd", x.data_ptr(), w.data_ptr(), y.data_ptr(), _fptr(scale), _fptr(shift), ACTS[act],
        B, Cin, out_channels, H, W, fh, fw, kh, kw, padding[0], padding[1], dilation[0], dilation[1], groups,
        PAD_MODES[padding_mode], _DTYPES[dt], layout, row, _stream())
    return y


def meta_conv2d(x, w, out_channels, kernel_size, padding=(0, 0), dilation=(1, 1), groups=1,
                padding_mode="zeros"):
    """Per-sample dynamic convolution, stride 1 (MetaConv2d semantics); w is (N, hyper_params)."""
    _require_cuda(x, w)
    _require_inference(x, w)
    dt = _compute_dtype(x)
    x = x.to(dt).contiguous()
    w = w.to(dt).reshape(w.shape[0], -1).contiguous()
    N, Cin, H, W = x.shape
    kh, kw = kernel_size
    if w.shape[0] != N or w.shape[1] != out_channels * (Cin // groups) * kh * kw:
        raise ValueError(f"weights {tuple(w.shape)} do not match input {tuple(x.shape)}")
    Ho = H + 2 * padding[0] - dilation[0] * (kh - 1)
    Wo = W + 2 * padding[1] - dilation[1] * (kw - 1)
    y = torch.empty((N, out_channels, Ho, Wo), dtype=dt, device=x.device)
    _call("hsb_meta_conv2d_fwd", x.data_ptr(), w.data_ptr(), y.data_ptr(), N, Cin, out_channels, H, W, kh, kw,
        padding[0], padding[1], dilation[0], dilation[1], groups, PAD_MODES[padding_mode], _DTYPES[dt],
        _stream())
    return y


def weights_to_patch_major(w):
    """(B, hp, fh, fw) contiguous -> same logical tensor in patch-major storage."""
    _require_cuda(w)
    if w.dtype not in _DTYPES:
        raise TypeError(f"unsupported dtype {w.dtype}")
    w = w.contiguous()
    B, hp, fh, fw = w.shape
    row = (hp + 7) // 8 * 8
    buf = torch.empty((B, fh, fw, row), dtype=w.dtype, device=w.device)
    _call("hsb_weights_to_patch_major", w.data_ptr(), buf.data_ptr(), B, hp, fh, fw, row, _DTYPES[w.dtype], _stream())
    return buf[..., :hp].permute(0, 3, 1, 2)


def decoder_input(coords, feature, prev):
    """cat(coords, feature, bilinear_upsample(prev, feature.shape[-2:])) in one kernel.

    ``coords`` is (1 or B, 2, H, W) (broadcast over the batch) or None, ``feature`` (B, Cf, H, W) in any memory
    format, ``prev`` (B, Cp, h, w) or None."""
    _require_cuda(feature, prev, coords)
    _require_inference(feature, prev)
    dt = _compute_dtype(feature)
    B, Cf, H, W = feature.shape
    if feature.dtype != dt:
        feature = feature.to(dt)
    Cc = Cp = 0
    h = w = 1
    cptr = pptr = None
    if coords is not None:
        coords = coords[:1].to(dt).contiguous()
        Cc, cptr = coords.shape[1], coords.data_ptr()
    if prev is not None:
        prev = prev.to(dt).contiguous()
        Cp, h, w, pptr = prev.shape[1], prev.shape[2], prev.shape[3], prev.data_ptr()
    out = torch.empty((B, Cc + Cf + Cp, H, W), dtype=dt, device=feature.device)
    st = feature.stride()
    _call("hsb_decoder_input_fwd", cptr, feature.data_ptr(), pptr, out.data_ptr(), B, Cc, Cf, Cp, H, W, h, w,
          st[0], st[1], st[2], st[3], _DTYPES[dt], _stream())
    return out


def upsample_argmax(logits, size):
    """uint8 label map = argmax over classes of the bilinearly upsampled logits (never materialised)."""
    _require_cuda(logits)
    dt = logits.dtype
    if dt not in _DTYPES:
        raise TypeError(f"unsupported dtype {dt}")
    logits = logits.contiguous()
    B, C, h, w = logits.shape
    H, W = size
    labels = torch.empty((B, H, W), dtype=torch.uint8, device=logits.device)
    _call("hsb_upsample_argmax_fwd", logits.data_ptr(), labels.data_ptr(), B, C, h, w, H, W, _DTYPES[dt], _stream())
    return labels


# ---------------------------------------------------------------------------------------------------------------------
# encoder epilogues (engine path): channels-last activations of the static encoder, BatchNorm already folded
# ---------------------------------------------------------------------------------------------------------------------
def nhwc_epilogue_ok(x: torch.Tensor) -> bool:
    """True when ``x`` is a CUDA (N, C, H, W) tensor stored channels-last whose rows the 16-byte kernels can walk."""
    return (x.is_cuda and x.dim() == 4 and x.dtype in _DTYPES and not _needs_grad(x)
            and (x.shape[1] * x.element_size()) % 16 == 0 and x.shape[1] * x.element_size() <= 16 * 1024
            and x.is_contiguous(memory_format=torch.channels_last) and x.data_ptr() % 16 == 0)


def bias_act_nhwc_(x, bias, act="none", residual=None, pool=False):
    """In place: x <- act(x + bias[c]) (+ residual) on a channels-last tensor; with ``pool`` also returns per-chunk
    sums of the result, (N, chunks, C) float32 -- the squeeze of squeeze-and-excitation before its division by H*W,
    in a fixed summation order (see :func:`pooled_mean`)."""
    if not nhwc_epilogue_ok(x):
        raise ValueError("bias_act_nhwc_ needs a channels-last CUDA tensor with 16-byte channel rows")
    N, C, H, W = x.shape
    if residual is not None and not (nhwc_epilogue_ok(residual) and residual.shape == x.shape and residual.dtype == x.dtype):
        raise ValueError("residual must match x (shape, dtype, channels-last)")
    if bias is not None and not (bias.dtype == torch.float32 and bias.numel() == C and bias.is_contiguous()):
        raise ValueError("bias must be a contiguous float32 vector of C elements")
    partial = None
    if pool:
        chunks = _lib.load().hsb_bias_act_nhwc_chunks(C, H * W, _DTYPES[x.dtype])
        if chunks <= 0:
            raise ValueError("bad geometry for the pooled epilogue")
        partial = torch.empty((N, chunks, C), dtype=torch.float32, device=x.device)
    _call("hsb_bias_act_nhwc_fwd", x.data_ptr(), _fptr(bias), _fptr(residual), x.data_ptr(), _fptr(partial),
          N, H * W, C, ACTS[act], _DTYPES[x.dtype], _stream())
    return (x, partial) if pool else x


def pooled_mean(partial, hw, dtype):
    """(N, chunks, C) partial sums -> F.adaptive_avg_pool2d result (N, C, 1, 1)."""
    N, _, C = partial.shape
    return partial.sum(dim=1).mul_(1.0 / hw).to(dtype).view(N, C, 1, 1)


def channel_gate_nhwc_(x, gate):
    """In place: x <- x * sigmoid(gate[n, c]); ``gate`` is (N, C, 1, 1) (the excitation of squeeze-and-excitation)."""
    if not nhwc_epilogue_ok(x):
        raise ValueError("channel_gate_nhwc_ needs a channels-last CUDA tensor with 16-byte channel rows")
    N, C, H, W = x.shape
    gate = gate.reshape(N, C).to(x.dtype).contiguous()
    _call("hsb_channel_gate_nhwc_fwd", x.data_ptr(), gate.data_ptr(), x.data_ptr(), N, H * W, C, _DTYPES[x.dtype], _stream())
    return x


# ---------------------------------------------------------------------------------------------------------------------
# training path: autograd Functions over the backward entry points (hyperseg_v0_1 / BASELINE config 4)
# ---------------------------------------------------------------------------------------------------------------------
class _PatchConvFn(torch.autograd.Function):
    """Patch-wise convolution with gradients for x and the per-patch weights (any kernel / groups / dilation / pad mode)."""

    @staticmethod
    def forward(ctx, x, w, out_channels, kernel_size, padding, dilation, groups, padding_mode):
        xp, wp, layout, row, dt = _prep_xw(x, w)
        if kernel_size == (1, 1) and padding == (0, 0):
            y = _patch_conv1x1_fwd(xp, wp, out_channels, groups)
        else:
            y = _patch_conv_fwd(xp, wp, out_channels, kernel_size, padding, dilation, groups, padding_mode)
        ctx.save_for_backward(xp, wp)
        ctx.meta = (out_channels, kernel_size, padding, dilation, groups, padding_mode, layout, row, dt)
        return y

    @staticmethod
    def backward(ctx, dy):
        xp, wp = ctx.saved_tensors
        out_channels, (kh, kw), (ph_, pw_), (dh, dw_), groups, mode, layout, row, dt = ctx.meta
        dy = dy.to(dt).contiguous()
        B, Cin, H, W = xp.shape
        hp, fh, fw = wp.shape[1], wp.shape[2], wp.shape[3]
        geom = (B, Cin, out_channels, H, W, fh, fw, kh, kw, ph_, pw_, dh, dw_, groups, PAD_MODES[mode], _DTYPES[dt], layout, row)
        dx = dwt = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(xp)
            _call("hsb_patch_conv_bwd_input", wp.data_ptr(), dy.data_ptr(), dx.data_ptr(), *geom, _stream())
        if ctx.needs_input_grad[1]:
            if layout == W_PATCH_MAJOR:
                buf = torch.zeros((B, fh, fw, row), dtype=dt, device=xp.device)
                dwt = buf[..., :hp].permute(0, 3, 1, 2)
            else:
                buf = dwt = torch.empty((B, hp, fh, fw), dtype=dt, device=xp.device)
            _call("hsb_patch_conv_bwd_weight", xp.data_ptr(), dy.data_ptr(), buf.data_ptr(), *geom, _stream())
        return dx, dwt, None, None, None, None, None, None


class _HeadFn(torch.autograd.Function):
    """signal -> per-patch weights head with gradients for the signal map and the static head weights."""

    @staticmethod
    def forward(ctx, s, ws, sig_index, sig_ch, hp, groups):
        dt = _compute_dtype(s)
        sp = s.detach().to(dt).contiguous()
        out = _signal2weights_fwd(sp, ws.detach(), sig_index, sig_ch, hp, groups)
        ctx.save_for_backward(sp, ws.detach().to(dt).reshape(ws.shape[0], -1).contiguous())
        ctx.meta = (sig_index, sig_ch, hp, groups, dt, tuple(ws.shape))
        return out

    @staticmethod
    def backward(ctx, dwout):
        sp, ws2 = ctx.saved_tensors
        sig_index, sig_ch, hp, groups, dt, ws_shape = ctx.meta
        B, C, fh, fw = sp.shape
        g, layout, row = weight_layout(dwout.to(dt))
        geom = (B, C, sig_index, sig_ch, ws2.shape[0], hp, groups, fh, fw, C * fh * fw, fh * fw, 1, _DTYPES[dt], layout, row)
        ds = dws = None
        if ctx.needs_input_grad[0]:
            ds = torch.empty_like(sp)
            _call("hsb_signal2weights_bwd_signal", ws2.data_ptr(), g.data_ptr(), ds.data_ptr(), *geom, _stream())
        if ctx.needs_input_grad[1]:
            dws = torch.empty_like(ws2)
            _call("hsb_signal2weights_bwd_weight", sp.data_ptr(), g.data_ptr(), dws.data_ptr(), *geom, _stream())
            dws = dws.reshape(ws_shape)
        return ds, dws, None, None, None, None
