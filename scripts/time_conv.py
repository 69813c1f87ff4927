"""Graph-timed per-patch 1x1 levels of HyperSeg-M at batch 8 (16x32 patches), bf16, patch-major weights, BN + ReLU fused;
checked against the general kernel first.  HSB_NO_RING=1 times the one-shot kernel instead of the persistent ring kernel.

    python scripts/time_conv.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from check_ir2 import bn, graph_time, rnd  # noqa: E402
from hyperseg_b200 import _lib, ops  # noqa: E402
from oracle import hyperseg_oracle as orc  # noqa: E402

DEV, B, PEAK = "cuda", 8, 6539.5
LEVELS = [(82, 64, 32), (94, 32, 16), (44, 16, 8)]
a = torch.randn(8192, 8192, device=DEV, dtype=torch.bfloat16)
for _ in range(30):
    a @ a
torch.cuda.synchronize()
tb = tt = 0.0
for lv, (cin, cout, div) in enumerate(LEVELS):
    h, w_ = 512 // div, 1024 // div
    sc, sh = bn(cout, 5)
    sc, sh = sc.to(DEV), sh.to(DEV)
    # parity at one image against the float64 oracle
    x1 = rnd((1, cin, h, w_), 1).to(DEV, torch.bfloat16)
    w1 = rnd((1, cin * cout, 16, 32), 2, 0.3).to(DEV, torch.bfloat16)
    ref = orc.patch_conv1x1(x1.float().cpu(), w1.float().cpu(), cout, 1, sc.cpu(), sh.cpu(), "relu")
    y1 = ops.patch_conv1x1(x1, ops.weights_to_patch_major(w1), cout, 1, sc, sh, "relu")
    kern = _lib.last_kernel()
    err = float((y1.float().cpu() - ref).abs().max() / ref.abs().max())
    sets = []
    for k in range(3):
        x = rnd((B, cin, h, w_), 10 + k).to(DEV, torch.bfloat16)
        wt = ops.weights_to_patch_major(rnd((B, cin * cout, 16, 32), 20 + k, 0.3).to(DEV, torch.bfloat16))
        sets.append((x, wt))
    us = graph_time([lambda x=x, wt=wt: ops.patch_conv1x1(x, wt, cout, 1, sc, sh, "relu") for x, wt in sets])
    nbytes = 2 * B * (cin * h * w_ + cin * cout * 512 + cout * h * w_)
    tb += nbytes; tt += us
    print(f"L{lv} conv1x1 {cin}->{cout} ({kern}): rel err {err:.2e}  {us:.1f} us  {nbytes / us * 1e-3:.0f} GB/s  frac {nbytes / us * 1e-3 / PEAK:.3f}", flush=True)
print(f"aggregate: {tt:.1f} us  frac {tb / tt * 1e-3 / PEAK:.3f}")
