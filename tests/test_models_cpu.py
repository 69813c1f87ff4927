"""Whole-model plumbing on CPU (BASELINE.json configs[0] and its siblings).

The mirror networks are built from hyperseg_b200.nn with seeded weights and executed with the dynamic ops
routed to the oracle; logits must equal the reference's (tests/golden/models.npz, produced by the unmodified
reference with the same seeded weights).  Also checks the drop-in contract: state_dict keys/shapes and the
head geometry (signal_index / channels / groups, incl. the reference's signal_index==0 quirk) equal the
reference's, and -- when /root/reference is present -- that a reference state_dict loads with strict=True.
"""
import importlib
import os
import sys
import types

import numpy as np
import pytest
import torch

import cases
from conftest import rel_err
from hyperseg_b200.synthetic import CONFIGS, build_model, deterministic_init, synthetic_frames
from oracle import hyperseg_oracle as orc

TOL = 5e-6


@pytest.mark.parametrize("name", sorted(cases.MODEL_CASES))
def test_model_logits_match_reference(name, golden_models):
    mc = cases.MODEL_CASES[name]
    model = build_model(mc["config"], seed=0)
    x = synthetic_frames(mc["B"], mc["H"], mc["W"])
    stats = golden_models[f"{name}/stats"]
    assert abs(x.double().sum().item() - stats[3]) < 1e-6, "seeded frames differ from the generator's"
    with torch.no_grad(), orc.use_oracle_ops():
        y = model(x)
    assert y.shape == (mc["B"], CONFIGS[mc["config"]]["num_classes"], mc["H"], mc["W"])
    st = cases.MODEL_STRIDE
    ref = torch.from_numpy(golden_models[f"{name}/logits"])
    assert rel_err(y[:, :, ::st, ::st], ref) < TOL
    assert abs(y.std().item() - stats[1]) < 1e-4 * stats[1] + 1e-6
    agree = (y.argmax(1).to(torch.uint8) == torch.from_numpy(golden_models[f"{name}/argmax"])).float().mean().item()
    assert agree > 0.9999


@pytest.mark.parametrize("config", sorted(CONFIGS))
def test_state_dict_contract(config, golden_divide):
    model = build_model(config)
    sd = model.state_dict()
    keys = sorted(sd)
    ref_keys = [str(k) for k in golden_divide[f"geometry/{config}/keys"]]
    assert keys == ref_keys
    ref_shapes = [str(s) for s in golden_divide[f"geometry/{config}/shapes"]]
    assert [str(tuple(sd[k].shape)) for k in keys] == ref_shapes
    heads = []
    for _, m in model.named_modules():
        conv = getattr(m, "signal2weights", None)
        if conv is not None:
            hp = getattr(m, "hyper_params", getattr(m, "target_params", -1))
            heads.append((int(m.signal_index), int(m.signal_channels), int(conv.groups), int(conv.out_channels), int(hp)))
    if hasattr(model.weight_mapper, "out_conv"):
        oc = model.weight_mapper.out_conv
        for i in range(len(oc.out_channels)):
            conv = getattr(oc, f"conv_{i}")
            heads.append((int(oc._ranges[i]), int(conv.in_channels), int(conv.groups), int(conv.out_channels), -1))
    assert np.array_equal(np.array(heads, dtype=np.int64), golden_divide[f"geometry/{config}"])


def test_v1_0_heads_all_read_from_signal_index_zero():
    """Reference quirk: init_signal2weights never propagates its offset back."""
    model = build_model("hyperseg-m")
    idx = [m.signal_index for m in model.modules() if getattr(m, "signal2weights", None) is not None]
    assert len(idx) == 5 and all(i == 0 for i in idx)
    uni = build_model("hyperseg-s-cityscapes")
    idx = [m.signal_index for m in uni.modules() if getattr(m, "signal2weights", None) is not None]
    assert idx == [0, 576, 704, 768]


def test_divide_feature_golden(golden_divide):
    from hyperseg_b200.nn.hyperseg_v0_1 import divide_feature_legacy
    from hyperseg_b200.nn.hyperseg_v1_0 import divide_feature
    for i, (inf, outs, unit) in enumerate(cases.DIVIDE_CASES):
        assert np.array_equal(divide_feature(inf, list(outs), min_unit=unit), golden_divide[f"v1_0/{i}"]), i
        legacy = golden_divide[f"legacy/{i}"]
        if legacy[0] >= 0:
            assert np.array_equal(divide_feature_legacy(inf, list(outs), unit), legacy), i


def test_forward_on_cpu_fails_loudly():
    """No CPU fallback: without the oracle patched in, a CPU forward must raise, not silently compute."""
    model = build_model("hyperseg-m")
    with torch.no_grad(), pytest.raises(RuntimeError, match="CUDA"):
        model(synthetic_frames(1, 64, 128))


def test_cpu_tensors_are_refused_under_autograd_too():
    from hyperseg_b200 import ops
    x = torch.zeros(1, 2, 2, 2, requires_grad=True)
    w = torch.zeros(1, 4, 1, 1)
    with pytest.raises((NotImplementedError, RuntimeError)):
        ops.patch_conv1x1(x, w, 2)


@pytest.mark.parametrize("fused", [True, False])
@pytest.mark.parametrize("config", ["hyperseg-m", "hyperseg-l-voc"])
def test_engine_batchnorm_folding_preserves_logits(config, fused):
    """What SegmentationEngine does to the static parts before it casts them: BatchNorm scale into the convolution,
    shift into the bias or (fused epilogues) a FoldedBatchNorm -- same logits as the unfolded model."""
    import copy
    from hyperseg_b200.engine import fold_static_batchnorms
    from hyperseg_b200.nn.efficientnet import FoldedBatchNorm
    from oracle.hyperseg_oracle import use_oracle_ops
    model = build_model(config).eval()
    folded = copy.deepcopy(model)
    n = fold_static_batchnorms(folded, fused_epilogues=fused)
    assert n > 60
    kept = [m for m in folded.backbone.modules() if isinstance(m, FoldedBatchNorm)]
    assert (len(kept) > 60) == fused
    assert not any(isinstance(m, torch.nn.BatchNorm2d) for m in folded.backbone.modules())
    x = synthetic_frames(1, 128, 128)
    with torch.no_grad(), use_oracle_ops():
        ref, got = model(x), folded(x)
    assert (got - ref).abs().max().item() < 2e-4 * ref.abs().max().item()


REFERENCE = "/root/reference"


@pytest.mark.skipif(not os.path.isdir(REFERENCE), reason="reference checkout not present (GPU box)")
@pytest.mark.parametrize("config", ["hyperseg-m", "hyperseg-s-cityscapes", "hyperseg-l-voc"])
def test_reference_state_dict_loads_strict_and_live_parity(config):
    sys.modules.setdefault("ffmpeg", types.ModuleType("ffmpeg"))
    if REFERENCE not in sys.path:
        sys.path.insert(0, REFERENCE)
    cfg = CONFIGS[config]
    mod = importlib.import_module("hyperseg.models." + cfg["module"])
    kwargs = {k: (list(v) if isinstance(v, list) else v) for k, v in cfg["kwargs"].items()}
    torch.manual_seed(3)
    ref = mod.hyperseg_efficientnet(cfg["model_name"], pretrained=False, num_classes=cfg["num_classes"], **kwargs)
    deterministic_init(ref, seed=7).eval()
    mine = build_model(config, seed=1)
    assert not mine.load_state_dict(ref.state_dict(), strict=True).missing_keys
    x = synthetic_frames(1, 128, 128, seed=9)
    with torch.no_grad():
        yr = ref(x)
        with orc.use_oracle_ops():
            ym = mine(x)
    assert rel_err(ym, yr) < TOL


@pytest.mark.skipif(not os.path.isdir(REFERENCE), reason="reference checkout not present (GPU box)")
def test_divide_feature_against_live_reference():
    sys.modules.setdefault("ffmpeg", types.ModuleType("ffmpeg"))
    if REFERENCE not in sys.path:
        sys.path.insert(0, REFERENCE)
    ref = importlib.import_module("hyperseg.models.hyperseg_v1_0")
    from hyperseg_b200.nn.hyperseg_v1_0 import divide_feature
    rng = np.random.RandomState(0)
    for _ in range(200):
        n = rng.randint(1, 7)
        unit = int(rng.choice([4, 8, 16, 32]))
        outs = [int(v) for v in rng.randint(1, 6000, size=n)]
        if rng.rand() < 0.4 and n > 1:
            outs[1] = outs[0]
        total = unit * int(rng.randint(n, 80))
        try:
            expect = ref.divide_feature(total, list(outs), min_unit=unit)
        except Exception:
            continue
        assert np.array_equal(divide_feature(total, list(outs), min_unit=unit), expect), (total, outs, unit)
