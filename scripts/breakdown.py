"""Where does a HyperSeg-M step go?  Section timings (CUDA events, eager, no graph) + encoder layout variants."""
import copy, json, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hyperseg_b200 import ops
from hyperseg_b200.engine import SegmentationEngine
from hyperseg_b200.synthetic import build_model, synthetic_frames
import torch.nn.functional as F

B, H, W = 8, 512, 1024
dev = "cuda"
torch.backends.cudnn.benchmark = True
model = build_model("hyperseg-m")
eng = SegmentationEngine(model, B, H, W, use_graph=False)
net = eng.net
x = synthetic_frames(B, H, W).to(dev).bfloat16()

def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

res = {}
with torch.no_grad():
    feats = net.backbone(x)
    sig = net.weight_mapper(feats[-1])
    res["backbone_nchw"] = timeit(lambda: net.backbone(x))
    res["weight_mapper"] = timeit(lambda: net.weight_mapper(feats[-1]))
    res["decoder_total"] = timeit(lambda: net.decoder([x] + feats[:-1], sig))
    logits = net.decoder([x] + feats[:-1], sig)
    res["argmax_u8"] = timeit(lambda: logits.argmax(1).to(torch.uint8))
    res["full_eager"] = timeit(lambda: net(x))
    # decoder glue alone (upsample + cats), level 4 sizes
    p3 = torch.randn(B, 16, 128, 256, device=dev, dtype=torch.bfloat16)
    f4 = feats[0]
    coords = net.decoder.get_image_coordinates(B, 256, 512, dev).to(torch.bfloat16)
    def glue():
        p = F.interpolate(p3, (256, 512), mode="bilinear", align_corners=False)
        p = torch.cat((f4, p), 1)
        return torch.cat([coords, p], 1)
    res["glue_L4"] = timeit(glue)
    p4 = torch.randn(B, 19, 256, 512, device=dev, dtype=torch.bfloat16)
    res["final_upsample"] = timeit(lambda: F.interpolate(p4, (512, 1024), mode="bilinear", align_corners=False))
    # channels_last encoder
    bb = copy.deepcopy(net.backbone).to(memory_format=torch.channels_last)
    xcl = x.contiguous(memory_format=torch.channels_last)
    res["backbone_channels_last"] = timeit(lambda: bb(xcl))
    res["backbone_cl_to_nchw_feats"] = timeit(lambda: [f.contiguous() for f in bb(xcl)])
    # graph replay of whole net
    eng2 = SegmentationEngine(model, B, H, W, use_graph=True)
    res["graph_step"] = timeit(lambda: eng2.step(), 20)
print(json.dumps({k: round(v, 3) for k, v in res.items()}))

# CPU oracle thread scaling (small frame)
from oracle import hyperseg_oracle as orc
xs = synthetic_frames(1, 128, 256)
out = {}
for nt in (8, 16, 32, 64, 128):
    if nt > (os.cpu_count() or 1): break
    torch.set_num_threads(nt)
    with torch.no_grad(), orc.use_oracle_ops(dtype=torch.float32):
        model(xs); t = time.perf_counter(); model(xs); model(xs); out[nt] = round((time.perf_counter() - t) / 2, 3)
print("cpu_oracle_128x256_s_per_frame", json.dumps(out))
