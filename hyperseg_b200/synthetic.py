"""Synthetic workloads: the reference's shipped model configurations, seeded weights and frames.

There are no datasets or checkpoints in the build/bench environment, so benchmarks and parity tests run
on random-initialised networks.  ``deterministic_init`` depends only on parameter *names and shapes*, so
the reference model and this package's mirror (identical state_dict keys) receive identical values.
"""
from __future__ import annotations

import zlib

import torch

# kwargs of the `model = partial(hyperseg_efficientnet, ...)` line of each reference config
CONFIGS = {
    # configs/train/cityscapes_efficientnet_b1_hyperseg-m.py:35-39
    "hyperseg-m": dict(
        module="hyperseg_v1_0", model_name="efficientnet-b1", num_classes=19, resolution=(512, 1024),
        kwargs=dict(levels=2, out_feat_scale=[1., 0.25, 0.25, 0.25, 0.25], kernel_sizes=[1, 1, 1, 3, 3],
                    level_channels=[64, 32, 16, 16, 16], expand_ratio=2, with_out_fc=False, decoder_dropout=None,
                    weight_groups=[32, 16, 8, 16, 4], decoder_groups=1, inference_hflip=True,
                    coords_res=[(512, 512), (512, 1024)])),
    # configs/train/cityscapes_efficientnet_b1_hyperseg-s.py:36-40
    "hyperseg-s-cityscapes": dict(
        module="hyperseg_v1_0_unify", model_name="efficientnet-b1", num_classes=19, resolution=(768, 1536),
        kwargs=dict(levels=2, out_feat_scale=[1., 0.166, 0.2, 0.25, 0.4], kernel_sizes=[1, 1, 1, 3, 3],
                    level_channels=[32, 16, 8, 8, 8], expand_ratio=2, with_out_fc=False, decoder_dropout=None,
                    weight_groups=[32, 16, 8, 16, 4], decoder_groups=1, inference_hflip=True, unify_level=4,
                    coords_res=[(768, 768), (768, 1536)])),
    # configs/train/camvid_efficientnet_b1_hyperseg-s.py:35-38
    "hyperseg-s-camvid": dict(
        module="hyperseg_v1_0", model_name="efficientnet-b1", num_classes=12, resolution=(576, 768),
        kwargs=dict(levels=2, kernel_sizes=(1, 1, 1, 3, 3), level_channels=[64, 32, 16, 16, 16], expand_ratio=2,
                    inference_hflip=True, with_out_fc=False, decoder_dropout=None, weight_groups=[64, 32, 32, 16, 8],
                    coords_res=[(576, 576), (576, 768)])),
    # configs/train/camvid_efficientnet_b1_hyperseg-l.py:35-38
    "hyperseg-l-camvid": dict(
        module="hyperseg_v1_0", model_name="efficientnet-b1", num_classes=12, resolution=(768, 1024),
        kwargs=dict(levels=2, kernel_sizes=(1, 1, 1, 3, 3, 3), level_channels=[64, 32, 16, 16, 16, 16],
                    expand_ratio=2, inference_hflip=True, with_out_fc=False, decoder_dropout=None,
                    weight_groups=[64, 32, 32, 16, 8, 8], coords_res=[(768, 768), (768, 1024)])),
    # configs/train/vocsbd_efficientnet_b3_hyperseg-l.py:32-34
    "hyperseg-l-voc": dict(
        module="hyperseg_v0_1", model_name="efficientnet-b3", num_classes=21, resolution=(512, 512),
        kwargs=dict(levels=3, kernel_sizes=(1, 1, 3, 3, 3, 3), expand_ratio=2, inference_hflip=True,
                    with_out_fc=False, decoder_dropout=None, weight_groups=16)),
}


def _name_seed(name: str, seed: int) -> int:
    return (zlib.crc32(name.encode()) + 1000003 * seed) % (2 ** 31 - 1)


@torch.no_grad()
def deterministic_init(model: torch.nn.Module, seed: int = 0) -> torch.nn.Module:
    """Fill every parameter and BatchNorm statistic with values that depend only on (name, shape, seed).

    Convolutions get a fan-in scaled normal (so activations stay O(1) through the network), BatchNorm gets
    non-trivial affine parameters and running statistics (so that fusing it wrongly is visible), and the
    signal->weights heads get a scale that makes the *generated* weights O(0.1)."""
    state = model.state_dict()
    for name in sorted(state):
        t = state[name]
        leaf = name.rsplit(".", 1)[-1]
        if leaf == "num_batches_tracked" or leaf.startswith("coord"):
            continue                       # cached coordinate grids are data, not weights
        g = torch.Generator().manual_seed(_name_seed(name, seed))
        if leaf == "running_var":
            v = torch.rand(t.shape, generator=g) + 0.5
        elif leaf == "running_mean":
            v = torch.randn(t.shape, generator=g) * 0.1
        elif t.dim() == 1 and leaf == "weight":
            v = torch.rand(t.shape, generator=g) * 0.4 + 0.8
        elif t.dim() == 1:
            v = torch.randn(t.shape, generator=g) * 0.1
        else:
            fan_in = t[0].numel()
            gain = 1.0
            v = torch.randn(t.shape, generator=g) * (gain / fan_in ** 0.5)
        t.copy_(v.to(t.dtype))
    return model


def synthetic_frames(batch: int, height: int, width: int, seed: int = 2, device="cpu", dtype=torch.float32):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(batch, 3, height, width, generator=g).to(device=device, dtype=dtype)


def build_model(config: str, seed: int = 0):
    """Instantiate this package's mirror of one of the reference configurations with seeded weights."""
    import importlib
    cfg = CONFIGS[config]
    mod = importlib.import_module(f"hyperseg_b200.nn.{cfg['module']}")
    kwargs = {k: (list(v) if isinstance(v, list) else v) for k, v in cfg["kwargs"].items()}
    model = mod.hyperseg_efficientnet(cfg["model_name"], pretrained=False, num_classes=cfg["num_classes"], **kwargs)
    return deterministic_init(model, seed).eval()
