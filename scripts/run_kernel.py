"""Launch one decoder kernel a few times at its HyperSeg-M batch-8 shape (for ncu captures)."""
import os, sys
os.environ.setdefault("HSB_VERBOSE", "1")
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hyperseg_b200 import ops

which = sys.argv[1] if len(sys.argv) > 1 else "ir"
B, dev, dt = 8, "cuda", torch.bfloat16
g = torch.Generator().manual_seed(0)
rnd = lambda *s, scale=1.0: (torch.randn(*s, generator=g) * scale).to(dev, dt)
bn = lambda n: ((torch.rand(n, generator=g) + 0.5).to(dev), (torch.randn(n, generator=g) * 0.1).to(dev))
if which == "ir":
    cin, hid, cout, h, w = 34, 68, 19, 256, 512
    x = rnd(B, cin, h, w); wt = ops.weights_to_patch_major(rnd(B, cin * hid + 9 * hid + hid * cout, 16, 32, scale=0.3))
    b1, b2, b3 = bn(hid), bn(hid), bn(cout)
    fn = lambda: ops.patch_ir(x, wt, hid, cout, b1, b2, b3)
elif which == "ir3":
    cin, hid, cout, h, w = 24, 48, 16, 128, 256
    x = rnd(B, cin, h, w); wt = ops.weights_to_patch_major(rnd(B, cin * hid + 9 * hid + hid * cout, 16, 32, scale=0.3))
    b1, b2, b3 = bn(hid), bn(hid), bn(cout)
    fn = lambda: ops.patch_ir(x, wt, hid, cout, b1, b2, b3)
elif which in ("ir2", "ir2_3"):
    cin, hid, cout, h, w = (34, 68, 19, 256, 512) if which == "ir2" else (24, 48, 16, 128, 256)
    x = rnd(B, cin, h, w); wt = ops.weights_to_patch_major(rnd(B, cin * hid + 9 * hid + hid * cout, 16, 32, scale=0.3))
    b1, b2, b3 = bn(hid), bn(hid), bn(cout)
    wa = ops.ir_arrange_weights(wt, cin, hid, cout, b1[0], b2[0], b3[0])
    fn = lambda: ops.patch_ir_arranged(x, wa, hid, cout, b1[1], b2[1], b3[1])
elif which == "conv0":
    x = rnd(B, 82, 16, 32); wt = ops.weights_to_patch_major(rnd(B, 82 * 64, 16, 32, scale=0.3)); sc, sh = bn(64)
    fn = lambda: ops.patch_conv1x1(x, wt, 64, 1, sc, sh, "relu")
elif which == "conv1":
    x = rnd(B, 94, 32, 64); wt = ops.weights_to_patch_major(rnd(B, 94 * 32, 16, 32, scale=0.3)); sc, sh = bn(32)
    fn = lambda: ops.patch_conv1x1(x, wt, 32, 1, sc, sh, "relu")
elif which == "conv2":
    x = rnd(B, 44, 64, 128); wt = ops.weights_to_patch_major(rnd(B, 44 * 16, 16, 32, scale=0.3)); sc, sh = bn(16)
    fn = lambda: ops.patch_conv1x1(x, wt, 16, 1, sc, sh, "relu")
elif which == "head4":
    s = rnd(B, 1280, 16, 32).abs(); ws = rnd(4216, 80, 1, 1, scale=0.2)
    fn = lambda: ops.signal2weights(s, ws, 0, 320, 4216, 4)
elif which == "head0":
    s = rnd(B, 1280, 16, 32).abs(); ws = rnd(5248, 13, 1, 1, scale=0.2)
    fn = lambda: ops.signal2weights(s, ws, 0, 416, 5248, 32)
elif which in ("head4a", "head3a"):
    cin, hid, cout = (34, 68, 19) if which == "head4a" else (24, 48, 16)
    hp = cin * hid + 9 * hid + hid * cout
    sig, grp = (320, 4) if which == "head4a" else (192, 16)       # HyperSeg-M: level-4 / level-3 slices of the 1280-channel signal
    s = rnd(B, 1280, 16, 32).abs(); ws = rnd(-(-hp // grp) * grp, sig // grp, 1, 1, scale=0.2)
    s1, s2, s3 = bn(hid)[0], bn(hid)[0], bn(cout)[0]
    hd = ops.ArrangedHead(ws, 0, sig, grp, 0, cin, hid, cout, s1, s2, s3)
    fn = lambda: ops.signal2weights_arranged(s, hd)
elif which == "epi":
    x = rnd(B, 96, 256, 512).contiguous(memory_format=torch.channels_last); sh = torch.randn(96, generator=g).to(dev)
    fn = lambda: ops.bias_act_nhwc_(x, sh, "silu", None, pool=True)
for _ in range(4):
    fn()
torch.cuda.synchronize()
print("done", which)
