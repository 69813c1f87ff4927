"""Round-2 experiment: the HSB_IR_X5D variant of the fused MetaBlock kernel fed with x in the blocked (B, H, W/8, C, 8)
layout (HSB_IR_XBLOCKED=1), checked against the oracle and timed.

    python -m hyperseg_b200.build --variant x5d HSB_IR_X5D
    HSB_LIBRARY=$PWD/hyperseg_b200/libhsb200_x5d.so HSB_IR_XBLOCKED=1 python scripts/x5d_blocked_check.py
"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hyperseg_b200 import _lib, ops
from oracle import hyperseg_oracle as orc

assert os.environ.get("HSB_IR_XBLOCKED") == "1" and "x5d" in os.environ.get("HSB_LIBRARY", ""), __doc__
dev, dt = "cuda", torch.bfloat16
g = torch.Generator().manual_seed(0)
bn = lambda n: ((torch.rand(n, generator=g) + 0.5), (torch.randn(n, generator=g) * 0.1))


def prepare(x, w, cout, bns):
    """x (B, C, H, W) bf16 on the device -> arguments of the C ABI with x in the blocked layout."""
    B, C, H, W = x.shape
    xb = x.view(B, C, H, W // 8, 8).permute(0, 2, 3, 1, 4).contiguous()
    wt = ops.weights_to_patch_major(w)
    _, layout, row = ops.weight_layout(wt)
    y = torch.empty(B, cout, H, W, dtype=dt, device=dev)
    aff = [t.to(dev, torch.float32).contiguous() for pair in bns for t in pair]
    return xb, wt, y, aff, layout, row, (B, C, H, W) + tuple(w.shape[-2:])


def launch(args, hid, cout):
    xb, wt, y, aff, layout, row, (B, C, H, W, fh, fw) = args
    st = _lib.load().hsb_patch_ir_fwd(xb.data_ptr(), wt.data_ptr(), y.data_ptr(), *[a.data_ptr() for a in aff], B, C, hid, cout,
                                      H, W, fh, fw, 0, 1, layout, row, torch.cuda.current_stream().cuda_stream)
    _lib.check(st, "hsb_patch_ir_fwd")
    return y


for (cin, hid, cout, ps, fh, fw) in [(34, 68, 19, 16, 2, 3), (24, 48, 16, 8, 3, 2)]:
    x = torch.randn(2, cin, fh * ps, fw * ps, generator=g).to(dt)
    w = (torch.randn(2, cin * hid + 9 * hid + hid * cout, fh, fw, generator=g) * 0.3).to(dt)
    bns = [bn(hid), bn(hid), bn(cout)]
    ref = orc.patch_ir(x.float(), w.float(), hid, cout, *bns)
    y = launch(prepare(x.to(dev), w.to(dev), cout, bns), hid, cout)
    err = (y.float().cpu() - ref).abs().max().item() / ref.abs().max().item()
    print(f"IR {cin}->{hid}->{cout} {ps}x{ps}: rel err {err:.2e} {'OK' if err < 1.5e-2 else 'FAIL'}")

# timing at the level-4 shape (B = 8), cold buffers
B, cin, hid, cout, H, W = 8, 34, 68, 19, 256, 512
sets = []
for _ in range(3):
    x = torch.randn(B, cin, H, W, generator=g).to(dev, dt)
    w = (torch.randn(B, cin * hid + 9 * hid + hid * cout, 16, 32, generator=g) * 0.3).to(dev, dt)
    sets.append(prepare(x, w, cout, [bn(hid), bn(hid), bn(cout)]))
for a in sets: launch(a, hid, cout)
torch.cuda.synchronize()
graph = torch.cuda.CUDAGraph(); side = torch.cuda.Stream()
with torch.cuda.stream(side):
    with torch.cuda.graph(graph, stream=side):
        for i in range(12): launch(sets[i % 3], hid, cout)
torch.cuda.synchronize()
ts = []
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(side):
        e0.record(side); graph.replay(); e1.record(side)
    side.synchronize(); ts.append(e0.elapsed_time(e1) / 12 * 1e3)
print("level-4 launch, x in the blocked layout:", sorted(ts)[2], "us (default kernel: 97.5 us)")
