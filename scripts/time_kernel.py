"""Device time of the decoder kernels at their HyperSeg-M batch-8 shapes (CUDA-graph replay over rotating cold
buffers), for A/B experiments:  HSB_LIBRARY=.../libhsb200_<variant>.so python scripts/time_kernel.py [names...]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hyperseg_b200 import ops

B, dev, dt = 8, "cuda", torch.bfloat16
g = torch.Generator().manual_seed(0)
rnd = lambda *s, scale=1.0: (torch.randn(*s, generator=g) * scale).to(dev, dt)
bn = lambda n: ((torch.rand(n, generator=g) + 0.5).to(dev), (torch.randn(n, generator=g) * 0.1).to(dev))
IR = {"ir": (34, 68, 19, 256, 512), "ir3": (24, 48, 16, 128, 256)}
CONV = {"conv0": (82, 64, 16, 32), "conv1": (94, 32, 32, 64), "conv2": (44, 16, 64, 128)}
HEAD = {"head0": (416, 32, 5248), "head1": (224, 16, 3008), "head2": (128, 8, 704), "head3": (192, 16, 2352), "head4": (320, 4, 4216)}


def make(which):
    if which in IR:
        cin, hid, cout, h, w = IR[which]
        x = rnd(B, cin, h, w); wt = ops.weights_to_patch_major(rnd(B, cin * hid + 9 * hid + hid * cout, 16, 32, scale=0.3))
        b1, b2, b3 = bn(hid), bn(hid), bn(cout)
        return lambda: ops.patch_ir(x, wt, hid, cout, b1, b2, b3)
    if which in CONV:
        cin, cout, h, w = CONV[which]
        x = rnd(B, cin, h, w); wt = ops.weights_to_patch_major(rnd(B, cin * cout, 16, 32, scale=0.3)); sc, sh = bn(cout)
        return lambda: ops.patch_conv1x1(x, wt, cout, 1, sc, sh, "relu")
    sc, groups, hp = HEAD[which]
    s = rnd(B, 1280, 16, 32).abs(); ws = rnd((hp + groups - 1) // groups * groups, sc // groups, 1, 1, scale=0.2)
    return lambda: ops.signal2weights(s, ws, 0, sc, hp, groups)


names = sys.argv[1:] or list(IR) + list(CONV) + list(HEAD)
out = []
with torch.no_grad():
    for which in names:
        fns = [make(which) for _ in range(3)]
        for f in fns: f()
        torch.cuda.synchronize()
        side = torch.cuda.Stream(); graph = torch.cuda.CUDAGraph(); iters = 12
        with torch.cuda.stream(side):
            for f in fns: f()
            side.synchronize()
            with torch.cuda.graph(graph, stream=side):
                for i in range(iters): fns[i % 3]()
        torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(side):
                e0.record(side); graph.replay(); e1.record(side)
            side.synchronize(); ts.append(e0.elapsed_time(e1) / iters)
        out.append(f"{which} {sorted(ts)[2] * 1e3:.1f}")
print(os.path.basename(os.environ.get("HSB_LIBRARY", "libhsb200.so")), "us/launch (median):", "  ".join(out))
