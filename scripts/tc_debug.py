"""Error anatomy of the tensor-core IR kernel vs the oracle (per patch / per channel / per tile position)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hyperseg_b200 import ops
from oracle import hyperseg_oracle as orc

torch.set_printoptions(linewidth=200, precision=3, sci_mode=False)
Cin, hid, Cout = 34, 68, 19
B, fh, fw = 1, 3, 4
g = torch.Generator().manual_seed(0)
x = torch.randn(B, Cin, fh * 16, fw * 16, generator=g).bfloat16()
w = (torch.randn(B, Cin * hid + 9 * hid + hid * Cout, fh, fw, generator=g) * 0.3).bfloat16()
bn = [(torch.rand(n, generator=g) + 0.5, torch.randn(n, generator=g) * 0.1) for n in (hid, hid, Cout)]
ref = orc.patch_ir(x.float(), w.float(), hid, Cout, *bn)
y = ops.patch_ir(x.cuda(), ops.weights_to_patch_major(w.cuda()), hid, Cout, *[(a.cuda(), b.cuda()) for a, b in bn])
torch.cuda.synchronize()
err = (y.float().cpu() - ref).abs()
print("max |ref|", ref.abs().max().item(), "max err", err.max().item(), "rel", (err.max() / ref.abs().max()).item())
print("finite:", torch.isfinite(y.float()).all().item())
pp = err.view(B, Cout, fh, 16, fw, 16).amax(dim=(1, 3, 5))
print("per-patch max err:\n", pp[0])
print("per-channel max err:", err.amax(dim=(0, 2, 3)))
pos = err.view(B, Cout, fh, 16, fw, 16).amax(dim=(0, 1, 2, 4))
print("per in-patch position max err (16x16):\n", pos)
print("sample y:", y[0, :4, 0, :6].float().cpu())
print("sample ref:", ref[0, :4, 0, :6])
