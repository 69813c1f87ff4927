"""One small launch of patch_ir2 (for compute-sanitizer): 34-68-19, 2 x (2 x 3) patches of 16 x 16."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hyperseg_b200 import ops
g = torch.Generator().manual_seed(0)
rnd = lambda *s, scale=1.0: (torch.randn(*s, generator=g) * scale).cuda().to(torch.bfloat16)
cin, hid, cout, ps = 34, 68, 19, 16
x = rnd(2, cin, 2 * ps, 3 * ps); wt = ops.weights_to_patch_major(rnd(2, cin * hid + 9 * hid + hid * cout, 2, 3, scale=0.3))
one = torch.ones(128, device="cuda"); zero = torch.zeros(128, device="cuda")
wa = ops.ir_arrange_weights(wt, cin, hid, cout, one[:hid], one[:hid], one[:cout])
y = ops.patch_ir_arranged(x, wa, hid, cout, zero[:hid], zero[:hid], zero[:cout])
torch.cuda.synchronize()
y0 = ops.patch_ir(x, wt, hid, cout, (one[:hid], zero[:hid]), (one[:hid], zero[:hid]), (one[:cout], zero[:cout]))
print("rel diff to round-1 kernel", float((y.float() - y0.float()).abs().max() / y0.float().abs().max()))
