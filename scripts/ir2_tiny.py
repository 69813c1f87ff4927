"""Small launches of patch_ir2 for compute-sanitizer: several patches per CTA (B x 16 x 32 patches over 148 CTAs)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hyperseg_b200 import ops
g = torch.Generator().manual_seed(0)
rnd = lambda *s, scale=1.0: (torch.randn(*s, generator=g) * scale).cuda().to(torch.bfloat16)
one = torch.ones(128, device="cuda"); zero = torch.zeros(128, device="cuda")
for cin, hid, cout, ps in ((24, 48, 16, 8), (34, 68, 19, 16)):
    x = rnd(1, cin, 16 * ps, 32 * ps); wt = ops.weights_to_patch_major(rnd(1, cin * hid + 9 * hid + hid * cout, 16, 32, scale=0.3))
    wa = ops.ir_arrange_weights(wt, cin, hid, cout, one[:hid], one[:hid], one[:cout])
    for rep in range(2):
        y = ops.patch_ir_arranged(x, wa, hid, cout, zero[:hid], zero[:hid], zero[:cout])
    torch.cuda.synchronize()
    y0 = ops.patch_ir(x, wt, hid, cout, (one[:hid], zero[:hid]), (one[:hid], zero[:hid]), (one[:cout], zero[:cout]))
    print((cin, hid, cout, ps), "rel diff to round-1 kernel", float((y.float() - y0.float()).abs().max() / y0.float().abs().max()), flush=True)
