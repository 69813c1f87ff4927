"""hyperseg_b200 -- B200-native (sm_100a) implementation of HyperSeg's decoder hot path.

Layout
  csrc/      hand-written CUDA kernels + the C ABI (include/hsb200.h) -> libhsb200.so
  _lib.py    ctypes binding of the C ABI (no fallback when the library is missing)
  ops.py     tensor-level wrappers (raw device pointers in, freshly allocated outputs out)
  nn/        host-side mirror of the reference's nn.Module surface (same class names, constructor
             arguments, forward signatures and state_dict keys as hyperseg/models/*)
  dist.py    batch sharding over one process per GPU (NCCL), logits all-gather / confusion-matrix all-reduce
"""
__version__ = "0.1.0"
