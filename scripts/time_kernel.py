"""Device time of one decoder kernel (CUDA-graph replay over rotating cold buffers), for A/B experiments."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hyperseg_b200 import ops
which = sys.argv[1]
B, dev, dt = 8, "cuda", torch.bfloat16
g = torch.Generator().manual_seed(0)
rnd = lambda *s, scale=1.0: (torch.randn(*s, generator=g) * scale).to(dev, dt)
bn = lambda n: ((torch.rand(n, generator=g) + 0.5).to(dev), (torch.randn(n, generator=g) * 0.1).to(dev))
fns = []
for _ in range(3):
    if which in ("ir", "ir3"):
        cin, hid, cout, h, w = (34, 68, 19, 256, 512) if which == "ir" else (24, 48, 16, 128, 256)
        x = rnd(B, cin, h, w); wt = ops.weights_to_patch_major(rnd(B, cin * hid + 9 * hid + hid * cout, 16, 32, scale=0.3))
        b1, b2, b3 = bn(hid), bn(hid), bn(cout)
        fns.append(lambda x=x, wt=wt, b1=b1, b2=b2, b3=b3: ops.patch_ir(x, wt, hid, cout, b1, b2, b3))
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
print(which, "HSB_IR_CTAS=" + os.environ.get("HSB_IR_CTAS", "-"), "us per launch:", [round(t * 1e3, 1) for t in sorted(ts)])
