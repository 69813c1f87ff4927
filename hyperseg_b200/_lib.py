"""ctypes binding of libhsb200.so (the C ABI in include/hsb200.h).

There is deliberately no fallback: if the shared library is missing or cannot be loaded, every
operator raises.  Build it with ``python -m hyperseg_b200.build`` (needs nvcc; cross-compiles for
sm_100a without a GPU).
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_int, c_int64, c_void_p, POINTER
from pathlib import Path

# HSB_LIBRARY selects an experiment build of the same library (hyperseg_b200.build --variant); never a fallback
LIB_PATH = Path(os.environ.get("HSB_LIBRARY") or (Path(__file__).resolve().parent / "libhsb200.so"))

HSB_F32, HSB_BF16 = 0, 1
ACT_NONE, ACT_RELU, ACT_RELU6, ACT_SILU = 0, 1, 2, 3
W_NCHW, W_PATCH_MAJOR = 0, 1
PAD_MODES = {"zeros": 0, "reflect": 1, "replicate": 2, "circular": 3}

_F = c_void_p  # float* passed as raw device address (None -> NULL)

# name -> argtypes, in the order of include/hsb200.h
SIGNATURES = {
    "hsb_version": [],
    "hsb_last_error": [],
    "hsb_last_kernel": [],
    "hsb_device_info": [POINTER(c_int), POINTER(c_int)],
    "hsb_patch_conv1x1_fwd": [c_void_p, c_void_p, c_void_p, _F, _F, c_int,
                              c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                              c_int, c_int, c_int64, c_void_p],
    "hsb_patch_ir_fwd": [c_void_p, c_void_p, c_void_p, _F, _F, _F, _F, _F, _F,
                         c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                         c_int, c_int, c_int, c_int64, c_void_p],
    "hsb_ir_arranged_row_elems": [c_int, c_int, c_int],
    "hsb_patch_ir_arranged_supported": [c_int, c_int, c_int, c_int],
    "hsb_ir_arrange_weights": [c_void_p, c_void_p, _F, _F, _F, c_int, c_int, c_int, c_int, c_int, c_int,
                               c_int, c_int, c_int64, c_void_p],
    "hsb_patch_ir_arranged_fwd": [c_void_p, c_void_p, c_void_p, _F, _F, _F,
                                  c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int64, c_void_p],
    "hsb_signal2weights_fwd": [c_void_p, c_void_p, c_void_p,
                               c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                               c_int64, c_int64, c_int64, c_int, c_int, c_int64, c_void_p],
    "hsb_head_packed_elems": [c_int, c_int, c_int],
    "hsb_head_pack": [c_void_p, c_void_p, _F, c_int, c_int, c_int, c_int, c_void_p],
    "hsb_signal2weights_packed_fwd": [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                      c_int64, c_int64, c_int64, c_void_p],
    "hsb_head_arranged_plan": [c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, POINTER(c_int64), POINTER(c_int), POINTER(c_int)],
    "hsb_head_pack_arranged": [c_void_p, c_void_p, c_void_p, _F, _F, _F, c_int, c_int, c_int, c_int, c_int,
                               c_int, c_int, c_int, c_int, c_void_p],
    "hsb_signal2weights_arranged_fwd": [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int,
                                        c_int, c_int, c_int64, c_int64, c_int64, c_void_p],
    "hsb_patch_conv_fwd": [c_void_p, c_void_p, c_void_p, _F, _F, c_int,
                           c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                           c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                           c_int, c_int, c_int, c_int64, c_void_p],
    "hsb_meta_conv2d_fwd": [c_void_p, c_void_p, c_void_p,
                            c_int, c_int, c_int, c_int, c_int,
                            c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                            c_int, c_int, c_void_p],
    "hsb_patch_conv_bwd_weight": [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                  c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int64, c_void_p],
    "hsb_patch_conv_bwd_input": [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                 c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int64, c_void_p],
    "hsb_signal2weights_bwd_signal": [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                      c_int64, c_int64, c_int64, c_int, c_int, c_int64, c_void_p],
    "hsb_signal2weights_bwd_weight": [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                      c_int64, c_int64, c_int64, c_int, c_int, c_int64, c_void_p],
    "hsb_decoder_input_fwd": [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                              c_int64, c_int64, c_int64, c_int64, c_int, c_void_p],
    "hsb_upsample_argmax_fwd": [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p],
    "hsb_weights_to_patch_major": [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int64, c_int, c_void_p],
    "hsb_bias_act_nhwc_chunks": [c_int, c_int64, c_int],
    "hsb_bias_act_nhwc_fwd": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int64, c_int, c_int, c_int, c_void_p],
    "hsb_channel_gate_nhwc_fwd": [c_void_p, c_void_p, c_void_p, c_int, c_int64, c_int, c_int, c_void_p],
}

_INT64_RESULTS = ("hsb_head_packed_elems", "hsb_ir_arranged_row_elems")

_lib = None


class HsbError(RuntimeError):
    """A libhsb200 entry point returned a negative status."""


def load() -> ctypes.CDLL:
    """Load libhsb200.so once; raise (never fall back) when it is not there."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise HsbError(
            f"{LIB_PATH} is missing: build it with `python -m hyperseg_b200.build`. "
            "hyperseg_b200 has no CPU or PyTorch fallback for the decoder hot path.")
    lib = ctypes.CDLL(str(LIB_PATH))
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here means header and library disagree
        fn.argtypes = argtypes
        fn.restype = c_char_p if name in ("hsb_last_error", "hsb_last_kernel") else (
            c_int64 if name in _INT64_RESULTS else c_int)
    _lib = lib
    return lib


def check(status: int, what: str) -> None:
    if status != 0:
        msg = load().hsb_last_error()
        raise HsbError(f"{what} failed with status {status}: {msg.decode() if msg else '?'}")


def last_kernel() -> str:
    """Name of the kernel the last entry point called from this thread launched (hsb_last_kernel)."""
    name = load().hsb_last_kernel()
    return name.decode() if name else ""


def version() -> int:
    return load().hsb_version()


def device_info() -> tuple[int, int]:
    sm, cc = c_int(0), c_int(0)
    check(load().hsb_device_info(ctypes.byref(sm), ctypes.byref(cc)), "hsb_device_info")
    return sm.value, cc.value
