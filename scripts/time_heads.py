"""Graph-timed weight heads of HyperSeg-M at batch 8 (16x32 signal map): reference-order rows for every level, arranged rows
for the two inverted-residual levels.  Same timing method as bench.py's `kernels` (12 launches per replay, 3 buffer sets).

    python scripts/time_heads.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from check_ir2 import graph_time, rnd  # noqa: E402
from hyperseg_b200 import ops  # noqa: E402

DEV, B, P = "cuda", 8, 16 * 32
PEAK = 6539.5
HEADS = [(416, 32, 5248, 5248), (224, 16, 3008, 3008), (128, 8, 704, 704), (192, 16, 2352, 2352), (320, 4, 4216, 4216)]
IR = {3: (24, 48, 16), 4: (34, 68, 19)}
a = torch.randn(8192, 8192, device=DEV, dtype=torch.bfloat16)
for _ in range(30):
    a @ a
torch.cuda.synchronize()
tot_b = tot_t = 0.0
for lv, (sc, g, och, hp) in enumerate(HEADS):
    sets = []
    for k in range(3):
        s = rnd((B, 1280, 16, 32), 40 + k).abs().to(DEV, torch.bfloat16)
        ws = rnd((och, sc // g, 1, 1), 50 + k, 0.2).to(DEV, torch.bfloat16)
        sets.append((s, ws))
    us = graph_time([lambda s=s, ws=ws: ops.signal2weights(s, ws, 0, sc, hp, g) for s, ws in sets])
    nbytes = 2 * (sc * P * B + hp * sc // g + hp * P * B)
    print(f"L{lv} head (reference order): {us:.1f} us  {nbytes / us * 1e-3:.0f} GB/s  frac {nbytes / us * 1e-3 / PEAK:.3f}", flush=True)
    if lv in IR:
        cin, hid, cout = IR[lv]
        one = torch.ones(128, device=DEV)
        heads = [ops.ArrangedHead(ws, 0, sc, g, 0, cin, hid, cout, one[:hid], one[:hid], one[:cout]) for _, ws in sets]
        us = graph_time([lambda s=s, h=h: ops.signal2weights_arranged(s, h) for (s, _), h in zip(sets, heads)])
        print(f"L{lv} head (arranged): {us:.1f} us  {nbytes / us * 1e-3:.0f} GB/s  frac {nbytes / us * 1e-3 / PEAK:.3f}", flush=True)
    tot_b += nbytes; tot_t += us
print(f"aggregate (arranged where available): {tot_t:.1f} us  frac {tot_b / tot_t * 1e-3 / PEAK:.3f}")
