"""Graph-timed level-input assembly (hsb_decoder_input_fwd) at the last two HyperSeg-M levels, batch 8, channels-last encoder
feature, exact 2x upsampling of the previous level.  HSB_NO_NHWC_GLUE=1 times the generic per-channel copy."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from check_ir2 import graph_time, rnd  # noqa: E402
from hyperseg_b200 import ops  # noqa: E402

DEV, B = "cuda", 8
for name, (cf, cp, H, W) in {"L4": (16, 16, 256, 512), "L3": (24, 16, 128, 256)}.items():
    coords = rnd((1, 2, H, W), 1).to(DEV, torch.bfloat16)
    sets = []
    for k in range(3):
        f = rnd((B, cf, H, W), 2 + k).to(DEV, torch.bfloat16).contiguous(memory_format=torch.channels_last)
        pv = rnd((B, cp, H // 2, W // 2), 5 + k).to(DEV, torch.bfloat16)
        sets.append((f, pv))
    for mode in ("1", "0"):
        os.environ["HSB_NO_NHWC_GLUE"] = mode
        us = graph_time([lambda f=f, pv=pv: ops.decoder_input(coords, f, pv) for f, pv in sets])
        nbytes = 2 * B * H * W * (cf + (2 + cf + cp)) + 2 * B * cp * H * W // 4
        print(f"{name} level input ({'generic' if mode == '1' else 'NHWC transposing copy'}): {us:.1f} us  {nbytes / us * 1e-3:.0f} GB/s", flush=True)

# label tail: 2x bilinear upsampling of the class logits fused with argmax (hsb_upsample_argmax_fwd)
sets = [rnd((B, 19, 256, 512), 30 + k).to(DEV, torch.bfloat16) for k in range(3)]
us = graph_time([lambda l=l: ops.upsample_argmax(l, (512, 1024)) for l in sets])
nb = 2 * B * 19 * 256 * 512 + B * 512 * 1024
print(f"label tail (19 x 256x512 -> 512x1024 labels): {us:.1f} us  {nb / us * 1e-3:.0f} GB/s", flush=True)
