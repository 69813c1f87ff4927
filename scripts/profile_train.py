"""Top CUDA kernels of one HyperSeg-L VOC training step (forward + backward + Adam, bf16 autocast), torch.profiler.

    python scripts/profile_train.py [batch]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hyperseg_b200 import _lib  # noqa: E402
from hyperseg_b200.synthetic import build_model, synthetic_frames  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
_lib.load()
model = build_model("hyperseg-l-voc", seed=0).cuda().train()
opt = torch.optim.Adam(model.parameters(), lr=1e-4, betas=(0.5, 0.999))
crit = torch.nn.CrossEntropyLoss(ignore_index=255)
x = synthetic_frames(B, 512, 512, seed=3).cuda()
y = torch.randint(0, 21, (B, 512, 512)).cuda()


def step():
    opt.zero_grad(set_to_none=True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        out = model(x)
    if out.shape[-2:] != y.shape[-2:]:
        out = torch.nn.functional.interpolate(out, y.shape[-2:], mode="bilinear", align_corners=False)
    loss = crit(out.float(), y)
    loss.backward()
    opt.step()
    return loss


for _ in range(3):
    step()
torch.cuda.synchronize()
from torch.profiler import ProfilerActivity, profile  # noqa: E402

with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(2):
        step()
    torch.cuda.synchronize()
rows = [(e.key, e.device_time_total / 2e3, e.count // 2) for e in prof.key_averages() if e.device_time_total > 0 and e.device_type.name == "CUDA"]
rows.sort(key=lambda r: -r[1])
total = sum(r[1] for r in rows)
print(f"device time per step: {total:.1f} ms over {sum(r[2] for r in rows)} launches")
for k, ms, n in rows[:40]:
    print(f"{ms:9.3f} ms  {n:5d}x  {k[:150]}")
