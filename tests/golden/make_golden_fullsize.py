"""Golden outputs of the UNMODIFIED reference at the benchmark sizes (BASELINE.json configs 2, 3 and 5), so that the path
bench.py times -- SegmentationEngine, bf16, folded BatchNorm, CUDA graph -- is pinned to the reference at the size it is
benchmarked at.  Run in the build container (imports /root/reference):

    python tests/golden/make_golden_fullsize.py      ->  tests/golden/fullsize.npz

Per configuration: one seeded frame (synthetic_frames(1, H, W, seed=FRAME_SEED)), reference fp32 logits at spatial
stride 8 (hyperseg/models/hyperseg_v1_0.py:71-91 forward), the full-resolution argmax map, and the top-2 logit margin at
full resolution quantised to uint8 (so that label agreement can be judged away from ties).
"""
import importlib
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.modules.setdefault("ffmpeg", types.ModuleType("ffmpeg"))
sys.path.insert(0, os.environ.get("HYPERSEG_REFERENCE", "/root/reference"))

from hyperseg_b200.synthetic import CONFIGS, deterministic_init, synthetic_frames  # noqa: E402

FULLSIZE_CASES = {"m_512x1024": ("hyperseg-m", 512, 1024), "s_city_768x1536": ("hyperseg-s-cityscapes", 768, 1536),
                  "s_camvid_576x768": ("hyperseg-s-camvid", 576, 768)}
FRAME_SEED, STRIDE = 1234, 8


def main():
    torch.set_grad_enabled(False)
    torch.set_num_threads(min(16, os.cpu_count() or 1))
    out = {}
    for name, (config, H, W) in FULLSIZE_CASES.items():
        cfg = CONFIGS[config]
        mod = importlib.import_module("hyperseg.models." + cfg["module"])
        kwargs = {k: (list(v) if isinstance(v, list) else v) for k, v in cfg["kwargs"].items()}
        model = mod.hyperseg_efficientnet(cfg["model_name"], pretrained=False, num_classes=cfg["num_classes"], **kwargs)
        deterministic_init(model, 0).eval()
        x = synthetic_frames(1, H, W, seed=FRAME_SEED)
        y = model(x)
        top2 = y.topk(2, dim=1).values
        margin = (top2[:, 0] - top2[:, 1]) / y.abs().max()
        out[f"{name}/logits"] = y[:, :, ::STRIDE, ::STRIDE].numpy()
        out[f"{name}/argmax"] = y.argmax(1).to(torch.uint8).numpy()
        out[f"{name}/margin_u8"] = (margin * 2550).clamp(0, 255).to(torch.uint8).numpy()     # 1 unit = 1/2550 of max |logit|
        out[f"{name}/stats"] = np.array([y.mean().item(), y.std().item(), y.abs().max().item(), x.double().sum().item()])
        print(f"{name}: logits {tuple(y.shape)} std {y.std():.4f} max {y.abs().max():.4f}")
    np.savez_compressed(os.path.join(HERE, "fullsize.npz"), **out)
    print("fullsize.npz", os.path.getsize(os.path.join(HERE, "fullsize.npz")) // 1024, "KiB")


if __name__ == "__main__":
    main()
