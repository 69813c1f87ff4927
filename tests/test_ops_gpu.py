"""Parity of the CUDA kernels (through the C ABI) with the reference -- run on the B200 with -m gpu.

Three kinds of evidence:
  * the committed golden vectors (outputs of the unmodified reference, tests/golden/ops.npz) in fp32;
  * the oracle on seeded inputs at shapes the oracle finishes in seconds, incl. the exact per-level shapes of
    HyperSeg-M at 512x1024 (one image), in fp32 and bf16, for every weight layout the ABI accepts;
  * size-independent properties at BASELINE.json's full size (batch 8): patch locality, batch independence,
    linearity in the weights, agreement between independent kernels.
Tolerances (relative to max |reference|): fp32 2e-5 (north_star allows 1e-3); bf16 I/O 1.5e-2 against the
float64 oracle evaluated on the same bf16-rounded inputs.
"""
import os

import pytest
import torch

import cases
from conftest import mirror_namespace, rel_err
from hyperseg_b200 import ops
from hyperseg_b200.synthetic import deterministic_init
from oracle import hyperseg_oracle as orc

pytestmark = pytest.mark.gpu
F32_TOL = 2e-5
BF16_TOL = 1.5e-2
DEV = "cuda"


def _rand(shape, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g) * scale


def _bn(n, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(n, generator=g) + 0.5, torch.randn(n, generator=g) * 0.1


# ---------------------------------------------------------------------------------------------------------
# golden vectors (reference outputs), fp32
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", sorted(cases.OP_CASES))
def test_golden_op_case_fp32(name, golden_ops):
    case = cases.OP_CASES[name]
    m = cases.build_op_module(mirror_namespace(), case)
    deterministic_init(m, cases.case_seed(name)).eval().to(DEV)
    x = torch.from_numpy(golden_ops[f"{name}/x"]).to(DEV)
    w = torch.from_numpy(golden_ops[f"{name}/w"]).to(DEV)
    with torch.no_grad():
        y = m(x, w)
    ref = torch.from_numpy(golden_ops[f"{name}/y"])
    assert y.shape == ref.shape and y.dtype == torch.float32
    assert rel_err(y.cpu(), ref) < F32_TOL


@pytest.mark.parametrize("name", sorted(cases.HEAD_CASES))
def test_golden_head_fp32(name, golden_ops):
    from hyperseg_b200.nn.hyperseg_v1_0 import HyperPatchNoPadding
    c = cases.HEAD_CASES[name]
    layer = HyperPatchNoPadding(c["hp"], 1, 1)
    layer.init_signal2weights(c["sc"], c["idx"], c["groups"])
    deterministic_init(layer, cases.case_seed(name)).to(DEV)
    s = torch.from_numpy(golden_ops[f"{name}/s"]).to(DEV)
    with torch.no_grad():
        w = layer.apply_signal2weights(s)
    assert w.shape == golden_ops[f"{name}/y"].shape
    assert w.stride(1) == 1, "heads must emit patch-major rows"
    assert rel_err(w.cpu(), golden_ops[f"{name}/y"]) < F32_TOL


@pytest.mark.parametrize("name", ["ir_L4", "ir_small", "nopad_p4", "block1x1", "mpblock_dw", "v01_ir"])
def test_golden_op_case_bf16(name, golden_ops):
    """bf16 I/O against the reference's fp32 answer: error budget = bf16 rounding of inputs and outputs."""
    case = cases.OP_CASES[name]
    m = cases.build_op_module(mirror_namespace(), case)
    deterministic_init(m, cases.case_seed(name)).eval().to(DEV)
    x = torch.from_numpy(golden_ops[f"{name}/x"]).to(DEV).bfloat16()
    w = torch.from_numpy(golden_ops[f"{name}/w"]).to(DEV).bfloat16()
    with torch.no_grad():
        y = m(x, w)
    assert y.dtype == torch.bfloat16
    assert rel_err(y.float().cpu(), golden_ops[f"{name}/y"]) < 3e-2


# ---------------------------------------------------------------------------------------------------------
# oracle on seeded inputs: HyperSeg-M level shapes (one image), every layout, fp32 + bf16
# ---------------------------------------------------------------------------------------------------------
M_1X1 = {"L0": (82, 64, 16, 32), "L1": (94, 32, 32, 64), "L2": (44, 16, 64, 128)}      # Cin, Cout, H, W (fh,fw = 16,32)
M_IR = {"L3": (24, 48, 16, 128, 256), "L4": (34, 68, 19, 256, 512)}                    # Cin, hid, Cout, H, W


def _layouts(w):
    """The same logical weights in the three storage forms the ABI accepts."""
    pm = ops.weights_to_patch_major(w)
    B, hp, fh, fw = w.shape
    wide = torch.zeros(B, fh, fw, hp + 24, device=w.device, dtype=w.dtype)
    wide[..., 5:5 + hp] = w.permute(0, 2, 3, 1)
    sliced = wide[..., 5:5 + hp].permute(0, 3, 1, 2)       # unaligned start, row stride > hp (unify-style slice)
    return {"nchw": w, "patch_major": pm, "sliced": sliced}


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("level", sorted(M_1X1))
def test_conv1x1_hyperseg_m_levels(level, dtype):
    Cin, Cout, H, W = M_1X1[level]
    x = _rand((1, Cin, H, W), 1).to(DEV, dtype)
    w = _rand((1, Cin * Cout, 16, 32), 2, 0.3).to(DEV, dtype)
    scale, shift = _bn(Cout, 3)
    ref = orc.patch_conv1x1(x.float().cpu(), w.float().cpu(), Cout, 1, scale, shift, "relu")
    tol = F32_TOL if dtype == torch.float32 else BF16_TOL
    for lname, wl in _layouts(w).items():
        y = ops.patch_conv1x1(x, wl, Cout, 1, scale.to(DEV), shift.to(DEV), "relu")
        assert rel_err(y.float().cpu(), ref) < tol, lname


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("level", sorted(M_IR))
def test_ir_hyperseg_m_levels(level, dtype):
    Cin, hid, Cout, H, W = M_IR[level]
    hp = Cin * hid + 9 * hid + hid * Cout
    x = _rand((1, Cin, H, W), 4).to(DEV, dtype)
    w = _rand((1, hp, 16, 32), 5, 0.3).to(DEV, dtype)
    bns = [_bn(hid, 6), _bn(hid, 7), _bn(Cout, 8)]
    ref = orc.patch_ir(x.float().cpu(), w.float().cpu(), hid, Cout, *bns)
    tol = F32_TOL if dtype == torch.float32 else BF16_TOL
    dbn = [(a.to(DEV), b.to(DEV)) for a, b in bns]
    for lname, wl in _layouts(w).items():
        y = ops.patch_ir(x, wl, hid, Cout, *dbn)
        assert rel_err(y.float().cpu(), ref) < tol, lname


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("geom", [(416, 32, 5248), (224, 16, 3008), (128, 8, 704), (192, 16, 2352), (320, 4, 4216)])
def test_heads_hyperseg_m(geom, dtype):
    sc, groups, hp = geom
    s = _rand((2, 1280, 16, 32), 9).abs().to(DEV, dtype)
    out_ch = -(-hp // groups) * groups
    ws = _rand((out_ch, sc // groups, 1, 1), 10, 0.2).to(DEV, dtype)
    ref = orc.signal2weights(s.float().cpu(), ws.float().cpu().flatten(1), 0, sc, hp, groups)
    tol = F32_TOL if dtype == torch.float32 else BF16_TOL
    y = ops.signal2weights(s, ws, 0, sc, hp, groups)
    assert y.shape == ref.shape
    assert rel_err(y.float().cpu(), ref) < tol
    y2 = ops.signal2weights(s.contiguous(memory_format=torch.channels_last), ws, 0, sc, hp, groups)
    assert rel_err(y2.float().cpu(), ref) < tol


def test_edge_shapes_fp32():
    """Ragged / degenerate geometry: odd channels, 1x1 and 1xN patch grids, patches of 1 pixel, residual."""
    for (B, Cin, Cout, H, W, fh, fw, g) in [(1, 3, 5, 7, 11, 7, 11, 1), (2, 9, 6, 6, 10, 3, 1, 3), (1, 1, 1, 4, 4, 1, 1, 1),
                                            (3, 16, 8, 2, 66, 1, 33, 1)]:
        x = _rand((B, Cin, H, W), 20).to(DEV)
        w = _rand((B, Cout * Cin // g, fh, fw), 21, 0.3).to(DEV)
        ref = orc.patch_conv1x1(x.cpu(), w.cpu(), Cout, g)
        assert rel_err(ops.patch_conv1x1(x, w, Cout, g).cpu(), ref) < F32_TOL
        assert rel_err(ops.patch_conv(x, w, Cout, (1, 1), (0, 0), groups=g).cpu(), ref) < F32_TOL
    for (B, Cin, hid, Cout, H, W, fh, fw) in [(1, 2, 2, 2, 2, 2, 1, 1), (2, 3, 7, 3, 9, 4, 3, 2), (1, 5, 10, 33, 8, 8, 2, 2),
                                               (1, 4, 8, 4, 48, 40, 2, 1)]:
        hp = Cin * hid + 9 * hid + hid * Cout
        x = _rand((B, Cin, H, W), 22).to(DEV)
        w = _rand((B, hp, fh, fw), 23, 0.3).to(DEV)
        bns = [_bn(hid, 24), _bn(hid, 25), _bn(Cout, 26)]
        res = Cin == Cout
        ref = orc.patch_ir(x.cpu(), w.cpu(), hid, Cout, *bns, residual=res)
        y = ops.patch_ir(x, w, hid, Cout, *[(a.to(DEV), b.to(DEV)) for a, b in bns], residual=res)
        assert rel_err(y.cpu(), ref) < F32_TOL


# ---------------------------------------------------------------------------------------------------------
# properties at BASELINE.json's full size (HyperSeg-M, batch 8, 512x1024)
# ---------------------------------------------------------------------------------------------------------
def _full_ir_inputs(dtype):
    Cin, hid, Cout, H, W = M_IR["L4"]
    hp = Cin * hid + 9 * hid + hid * Cout
    x = _rand((8, Cin, H, W), 30).to(DEV, dtype)
    w = ops.weights_to_patch_major(_rand((8, hp, 16, 32), 31, 0.3).to(DEV, dtype))
    bns = [tuple(t.to(DEV) for t in _bn(n, 32 + i)) for i, n in enumerate((hid, hid, Cout))]
    return x, w, bns, hid, Cout


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_full_size_ir_batch_independence_and_locality(dtype):
    x, w, bns, hid, Cout = _full_ir_inputs(dtype)
    y = ops.patch_ir(x, w, hid, Cout, *bns)
    assert torch.isfinite(y.float()).all()
    # an image's result does not depend on its neighbours in the batch, and equals the oracle's
    y3 = ops.patch_ir(x[3:4].contiguous(), w[3:4], hid, Cout, *bns)
    assert torch.equal(y3, y[3:4])
    ref = orc.patch_ir(x[3:4].float().cpu(), w[3:4].float().cpu(), hid, Cout, *[(a.cpu(), b.cpu()) for a, b in bns])
    assert rel_err(y3.float().cpu(), ref) < (F32_TOL if dtype == torch.float32 else BF16_TOL)
    # changing one patch's weights changes that patch's 16x16 outputs and nothing else
    w2 = w.clone()
    w2[5, :, 7, 9] *= 1.5
    y2 = ops.patch_ir(x, w2, hid, Cout, *bns)
    diff = (y2 != y)
    assert diff[5, :, 7 * 16:8 * 16, 9 * 16:10 * 16].any()
    diff[5, :, 7 * 16:8 * 16, 9 * 16:10 * 16] = False
    assert not diff.any()
    # run-to-run determinism
    assert torch.equal(ops.patch_ir(x, w, hid, Cout, *bns), y)


def test_full_size_conv1x1_linearity_and_kernel_agreement():
    Cin, Cout, H, W = M_1X1["L2"]
    x = _rand((8, Cin, H, W), 40).to(DEV)
    w1 = _rand((8, Cin * Cout, 16, 32), 41, 0.3).to(DEV)
    w2 = _rand((8, Cin * Cout, 16, 32), 42, 0.3).to(DEV)
    ya, yb, yab = (ops.patch_conv1x1(x, w, Cout) for w in (w1, w2, w1 + w2))
    assert rel_err((ya + yb).cpu(), yab.cpu()) < 1e-5
    # the dedicated 1x1 kernel and the generic kernel are independent implementations
    assert rel_err(ops.patch_conv(x, w1, Cout, (1, 1), (0, 0)).cpu(), ya.cpu()) < 1e-5
    # the fused epilogue equals a separate affine + relu
    scale, shift = (t.to(DEV) for t in _bn(Cout, 43))
    fused = ops.patch_conv1x1(x, w1, Cout, 1, scale, shift, "relu")
    manual = torch.relu(ya * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1))
    assert rel_err(fused.cpu(), manual.cpu()) < 1e-6


def test_full_size_head_then_conv_equals_oracle_on_one_image():
    """Head + 1x1 conv chained at level-0 size, the way the decoder runs them."""
    sc, groups, hp, Cin, Cout = 416, 32, 5248, 82, 64
    s = _rand((8, 1280, 16, 32), 50).abs().to(DEV)
    ws = _rand((hp, sc // groups, 1, 1), 51, 0.2).to(DEV)
    x = _rand((8, Cin, 16, 32), 52).to(DEV)
    wgt = ops.signal2weights(s, ws, 0, sc, hp, groups)
    y = ops.patch_conv1x1(x, wgt, Cout)
    wref = orc.signal2weights(s[6:7].cpu(), ws.cpu().flatten(1), 0, sc, hp, groups)
    yref = orc.patch_conv1x1(x[6:7].cpu(), wref, Cout)
    assert rel_err(y[6:7].cpu(), yref) < F32_TOL


# ---------------------------------------------------------------------------------------------------------
# host-side behaviour on a GPU box
# ---------------------------------------------------------------------------------------------------------
def test_cpu_tensors_are_refused():
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.patch_conv1x1(torch.zeros(1, 2, 2, 2), torch.zeros(1, 4, 1, 1), 2)


def test_abi_errors_surface_as_exceptions():
    from hyperseg_b200._lib import HsbError
    x = torch.zeros(1, 4, 6, 6, device=DEV)
    w = torch.zeros(1, 16, 4, 2, device=DEV)       # 6 % 4 != 0
    with pytest.raises(HsbError, match="divisible"):
        ops.patch_conv1x1(x, w, 4)


def test_kernels_follow_the_current_stream():
    x = _rand((2, 8, 8, 8), 60).to(DEV)
    w = _rand((2, 32, 2, 2), 61, 0.3).to(DEV)
    ref = ops.patch_conv1x1(x, w, 4)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        y = ops.patch_conv1x1(x, w, 4)
    side.synchronize()
    assert torch.equal(y, ref)


# ---------------------------------------------------------------------------------------------------------
# tensor-core (tcgen05) path of the fused inverted-residual block: bf16, 16x16 patches, patch-major weights
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("shape", [(34, 68, 19, 16), (26, 52, 19, 16), (22, 44, 12, 16), (24, 48, 16, 8), (14, 28, 8, 8)])
@pytest.mark.parametrize("grid", [(2, 3, 5), (1, 1, 1), (3, 1, 4), (1, 2, 1)])
def test_ir_tensor_core_path(shape, grid):
    """Every border case of the halo (corner / edge / interior / single patch) on the shipped IR level shapes."""
    Cin, hid, Cout, ps = shape
    B, fh, fw = grid
    hp = Cin * hid + 9 * hid + hid * Cout
    x = _rand((B, Cin, fh * ps, fw * ps), 70).to(DEV, torch.bfloat16)
    w = _rand((B, hp, fh, fw), 71, 0.3).to(DEV, torch.bfloat16)
    bns = [_bn(hid, 72), _bn(hid, 73), _bn(Cout, 74)]
    ref = orc.patch_ir(x.float().cpu(), w.float().cpu(), hid, Cout, *bns)
    y = ops.patch_ir(x, ops.weights_to_patch_major(w), hid, Cout, *[(a.to(DEV), b.to(DEV)) for a, b in bns])
    assert rel_err(y.float().cpu(), ref) < BF16_TOL


def test_ir_tensor_core_many_patches_per_cta():
    """More patches than CTAs: exercises the persistent loop, barrier phases and the prefetch ring."""
    Cin, hid, Cout, ps = 24, 48, 16, 8
    B, fh, fw = 4, 16, 24            # 1536 patches
    hp = Cin * hid + 9 * hid + hid * Cout
    x = _rand((B, Cin, fh * ps, fw * ps), 80).to(DEV, torch.bfloat16)
    w = _rand((B, hp, fh, fw), 81, 0.3).to(DEV, torch.bfloat16)
    bns = [_bn(hid, 82), _bn(hid, 83), _bn(Cout, 84)]
    ref = orc.patch_ir(x.float().cpu(), w.float().cpu(), hid, Cout, *bns)
    y = ops.patch_ir(x, ops.weights_to_patch_major(w), hid, Cout, *[(a.to(DEV), b.to(DEV)) for a, b in bns])
    assert rel_err(y.float().cpu(), ref) < BF16_TOL


# ---------------------------------------------------------------------------------------------------------
# decoder glue kernels against the stock PyTorch sequence they replace
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("geom", [(2, 16, 16, 128, 256, 64, 128), (1, 3, 7, 37, 52, 19, 26), (2, 10, 0, 16, 32, 1, 1),
                                  (1, 6, 5, 24, 40, 24, 40), (3, 24, 8, 32, 64, 16, 32), (1, 40, 32, 16, 24, 8, 12)])
def test_decoder_input_matches_interpolate_cat(geom, dtype):
    import torch.nn.functional as F
    B, Cf, Cp, H, W, h, w = geom
    feat = _rand((B, Cf, H, W), 90).to(DEV, dtype)
    prev = _rand((B, Cp, h, w), 91).to(DEV, dtype) if Cp else None
    xs, ys = torch.linspace(-1, 1, W), torch.linspace(-1, 1, H)
    coords = torch.stack((xs.view(1, W).expand(H, W), ys.view(H, 1).expand(H, W)), 0).unsqueeze(0).to(DEV, dtype)
    parts = [coords.expand(B, -1, -1, -1), feat]
    if Cp:
        parts.append(F.interpolate(prev, (H, W), mode="bilinear", align_corners=False) if (h, w) != (H, W) else prev)
    ref = torch.cat(parts, 1)
    for f in (feat, feat.contiguous(memory_format=torch.channels_last)):
        out = ops.decoder_input(coords, f, prev)
        assert out.shape == ref.shape and out.dtype == dtype and out.is_contiguous()
        tol = 1e-6 if dtype == torch.float32 else 8e-3          # one bf16 ulp on the interpolated channels
        assert (out.float() - ref.float()).abs().max().item() <= tol * max(1.0, ref.float().abs().max().item())
        assert torch.equal(out[:, :2 + Cf], ref[:, :2 + Cf])     # copied channels are bit-exact


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("geom", [(2, 19, 64, 128, 128, 256), (1, 12, 33, 20, 66, 40), (1, 5, 16, 16, 16, 16), (1, 21, 9, 7, 31, 23)])
def test_upsample_argmax_matches_interpolate_argmax(geom, dtype):
    import torch.nn.functional as F
    B, C, h, w, H, W = geom
    logits = _rand((B, C, h, w), 95).to(DEV, dtype)
    full = F.interpolate(logits, (H, W), mode="bilinear", align_corners=False) if (h, w) != (H, W) else logits
    ref = full.argmax(1)
    labels = ops.upsample_argmax(logits, (H, W))
    assert labels.dtype == torch.uint8 and labels.shape == (B, H, W)
    agree = (labels.long() == ref).float().mean().item()
    # ties / last-ulp differences of the interpolation may flip a label between two near-equal classes
    top2 = full.float().topk(2, dim=1).values
    margin = (top2[:, 0] - top2[:, 1])
    hard = (labels.long() != ref) & (margin > (1e-5 if dtype == torch.float32 else 4e-2))
    assert agree > 0.99 and not hard.any()


def test_engine_labels_match_model_argmax():
    from hyperseg_b200.engine import SegmentationEngine
    from hyperseg_b200.synthetic import build_model, synthetic_frames
    model = build_model("hyperseg-m", seed=0)
    frames = synthetic_frames(2, 64, 128).pin_memory()
    for graph in (False, True):
        eng = SegmentationEngine(model, 2, 64, 128, use_graph=graph)
        labels = eng(frames).clone()
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
            ref = model.cuda()(frames.cuda()).argmax(1).cpu()
        agree = (labels.long() == ref).float().mean().item()
        assert agree > 0.97, agree                        # bf16 engine vs autocast reference path
        assert eng.launches_per_step >= 16                # 5 heads + 5 patch kernels + 5 glue + tail, all ours


# ---- encoder epilogues (engine path) ---------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("geom", [(2, 96, 64, 128), (1, 16, 33, 20), (3, 1152, 4, 8), (2, 40, 1, 1), (1, 8, 7, 300)])
@pytest.mark.parametrize("act", ["silu", "none"])
def test_bias_act_nhwc_matches_torch(geom, dtype, act):
    """y = act(x + shift) (+ skip) and the squeeze-and-excitation mean, against the stock elementwise sequence
    (reference efficientnet.py:97-122 with the BatchNorm folded)."""
    from hyperseg_b200 import ops
    N, C, H, W = geom
    g = torch.Generator().manual_seed(C * 131 + H)
    x = (torch.randn(N, C, H, W, generator=g) * 2).to("cuda", dtype).contiguous(memory_format=torch.channels_last)
    res = torch.randn(N, C, H, W, generator=g).to("cuda", dtype).contiguous(memory_format=torch.channels_last)
    shift = torch.randn(C, generator=g).cuda()
    ref = x.float() + shift.view(1, -1, 1, 1)
    ref = torch.nn.functional.silu(ref) if act == "silu" else ref
    tol = 2e-6 if dtype == torch.float32 else 8e-3        # fp32: SFU exp / reciprocal approximations (~2e-7 relative)
    y, partial = ops.bias_act_nhwc_(x.clone(memory_format=torch.channels_last), shift, act, None, pool=True)
    assert partial.shape[0] == N and partial.shape[2] == C and partial.shape[1] <= 64
    mean = ops.pooled_mean(partial, H * W, dtype)
    assert y.is_contiguous(memory_format=torch.channels_last)
    assert (y.float() - ref).abs().max().item() <= tol * max(1.0, ref.abs().max().item())
    assert mean.shape == (N, C, 1, 1)
    assert (mean.float() - y.float().mean((2, 3), keepdim=True)).abs().max().item() <= tol * max(1.0, mean.float().abs().max().item())
    y2 = ops.bias_act_nhwc_(x.clone(memory_format=torch.channels_last), shift, act, res)
    assert (y2.float() - (ref + res.float())).abs().max().item() <= tol * max(1.0, ref.abs().max().item() + 4)
    # deterministic (fixed summation order)
    _, partial2 = ops.bias_act_nhwc_(x.clone(memory_format=torch.channels_last), shift, act, None, pool=True)
    assert torch.equal(partial, partial2)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("geom", [(2, 96, 64, 128), (1, 16, 33, 20), (3, 1152, 4, 8)])
def test_channel_gate_nhwc_matches_torch(geom, dtype):
    from hyperseg_b200 import ops
    N, C, H, W = geom
    g = torch.Generator().manual_seed(C + 7)
    x = torch.randn(N, C, H, W, generator=g).to("cuda", dtype).contiguous(memory_format=torch.channels_last)
    gate = (torch.randn(N, C, 1, 1, generator=g) * 3).to("cuda", dtype)
    ref = torch.sigmoid(gate) * x
    y = ops.channel_gate_nhwc_(x.clone(memory_format=torch.channels_last), gate)
    tol = 2e-6 if dtype == torch.float32 else 8e-3
    assert (y.float() - ref.float()).abs().max().item() <= tol * max(1.0, ref.float().abs().max().item())


@pytest.mark.gpu
def test_encoder_epilogues_refuse_other_layouts():
    from hyperseg_b200 import ops
    x = torch.zeros(1, 16, 4, 4, device="cuda")                       # NCHW-contiguous
    with pytest.raises(ValueError):
        ops.bias_act_nhwc_(x, torch.zeros(16, device="cuda"), "silu")
    x = torch.zeros(1, 3, 4, 4, device="cuda").contiguous(memory_format=torch.channels_last)   # 12-byte rows
    with pytest.raises(ValueError):
        ops.channel_gate_nhwc_(x, torch.zeros(1, 3, 1, 1, device="cuda"))


@pytest.mark.gpu
def test_engine_fused_encoder_epilogues_match_stock_encoder():
    """The engine with the one-pass encoder epilogues against the same engine on stock elementwise kernels."""
    from hyperseg_b200.engine import SegmentationEngine, fold_static_batchnorms
    from hyperseg_b200.nn.efficientnet import FoldedBatchNorm
    from hyperseg_b200.synthetic import build_model, synthetic_frames
    model = build_model("hyperseg-m", seed=0).eval()
    frames = synthetic_frames(2, 128, 256).pin_memory()
    eng = SegmentationEngine(model, batch=2, height=128, width=256, use_graph=False)
    assert any(isinstance(m, FoldedBatchNorm) and m.shift32 is not None for m in eng.net.modules())
    labels = eng(frames).clone()
    logits = eng.full_logits().float()
    for m in eng.net.modules():                      # same engine, stock path: y + shift, silu, mean, sigmoid * y
        if isinstance(m, FoldedBatchNorm):
            m.shift32 = None
    labels_stock = eng(frames).clone()
    logits_stock = eng.full_logits().float()
    assert (logits - logits_stock).abs().max().item() < 3e-2 * logits_stock.abs().max().item()
    assert (labels == labels_stock).float().mean().item() > 0.97


@pytest.mark.gpu
def test_engine_pipelined_stream_equals_blocking_calls():
    """submit()/collect() (upload of batch k+1 overlapping the forward of batch k) returns, in order, exactly the label
    maps of the blocking __call__ on the same batches."""
    from hyperseg_b200.engine import SegmentationEngine
    from hyperseg_b200.synthetic import build_model, synthetic_frames
    model = build_model("hyperseg-m", seed=0).eval()
    eng = SegmentationEngine(model, batch=2, height=128, width=256)
    batches = [synthetic_frames(2, 128, 256, seed=40 + i).pin_memory() for i in range(5)]
    want = [eng(b).clone() for b in batches]
    assert not torch.equal(want[0], want[1])
    got = []
    for k, b in enumerate(batches):
        eng.submit(b)
        if k > 0:
            got.append(eng.collect().clone())
    got.append(eng.collect().clone())
    assert len(got) == len(want) and all(torch.equal(g, w) for g, w in zip(got, want))
    with pytest.raises(RuntimeError):
        eng.collect()
    eng.submit(batches[0]); eng.submit(batches[1])
    with pytest.raises(RuntimeError):
        eng.submit(batches[2])
    assert torch.equal(eng.collect(), want[0]) and torch.equal(eng.collect(), want[1])


@pytest.mark.gpu
def test_engine_uint8_frames_equal_host_normalised_frames():
    """input_dtype=uint8: raw 8-bit frames normalised on the device give the label maps of the float32 path fed with the
    same frames normalised on the host (the reference's ToTensor + Normalize); wrong dtypes are refused."""
    from hyperseg_b200.engine import SegmentationEngine
    from hyperseg_b200.synthetic import build_model
    model = build_model("hyperseg-m", seed=0).eval()
    g = torch.Generator().manual_seed(77)
    u8 = torch.randint(0, 256, (2, 3, 128, 256), generator=g, dtype=torch.uint8).pin_memory()
    mean = torch.tensor([0.485, 0.456, 0.406]).view(1, 3, 1, 1)
    std = torch.tensor([0.229, 0.224, 0.225]).view(1, 3, 1, 1)
    f32 = ((u8.float() / 255 - mean) / std).pin_memory()
    e8 = SegmentationEngine(model, 2, 128, 256, input_dtype=torch.uint8)
    ef = SegmentationEngine(model, 2, 128, 256)
    a, b = e8(u8).clone(), ef(f32).clone()
    la, lb = e8.full_logits().float(), ef.full_logits().float()
    # same network input up to one bf16 rounding of (x * scale + bias) vs ((x / 255 - mean) / std)
    assert rel_err(la.cpu(), lb.cpu()) < 2e-2
    assert (a == b).float().mean().item() > 0.99
    with pytest.raises(ValueError):
        e8(f32)
    with pytest.raises(ValueError):
        ef.submit(u8)


# ---------------------------------------------------------------------------------------------------------
# restage-free tensor-core MetaBlock kernel (hsb_patch_ir_arranged_fwd) and the arranged weight head
# ---------------------------------------------------------------------------------------------------------
IR_SHAPES = [(34, 68, 19, 16), (26, 52, 19, 16), (22, 44, 12, 16), (24, 48, 16, 8), (14, 28, 8, 8)]
IR2_TOL = 1.2e-2      # measured on B200: <= 8.2e-3 of max |y| over every shape / border case below


def _run_arranged(x, w, hid, cout, bns):
    from hyperseg_b200 import _lib
    dev = [(a.to(DEV), b.to(DEV)) for a, b in bns]
    wa = ops.ir_arrange_weights(w, x.shape[1], hid, cout, dev[0][0], dev[1][0], dev[2][0])
    y = ops.patch_ir_arranged(x, wa, hid, cout, dev[0][1], dev[1][1], dev[2][1])
    assert _lib.last_kernel() == "patch_ir2_kernel"
    return y


@pytest.mark.parametrize("shape", IR_SHAPES)
@pytest.mark.parametrize("grid", [(2, 3, 5), (1, 1, 1), (3, 1, 4), (1, 2, 1)])
def test_ir_arranged_kernel_every_border_case(shape, grid):
    """Corner / edge / interior / single patches, rows and columns of one patch: mirror rows, mirror columns, halo columns
    from the neighbour tiles and from global memory (ends of a CTA's run)."""
    Cin, hid, Cout, ps = shape
    B, fh, fw = grid
    hp = Cin * hid + 9 * hid + hid * Cout
    x = _rand((B, Cin, fh * ps, fw * ps), 70).to(DEV, torch.bfloat16)
    w = _rand((B, hp, fh, fw), 71, 0.3).to(DEV, torch.bfloat16)
    bns = [_bn(hid, 72), _bn(hid, 73), _bn(Cout, 74)]
    ref = orc.patch_ir(x.float().cpu(), w.float().cpu(), hid, Cout, *bns)
    assert rel_err(_run_arranged(x, w, hid, Cout, bns).float().cpu(), ref) < IR2_TOL


@pytest.mark.parametrize("shape,grid", [((34, 68, 19, 16), (2, 16, 24)), ((24, 48, 16, 8), (4, 16, 24)), ((26, 52, 19, 16), (1, 13, 12)),
                                        ((24, 48, 16, 8), (3, 16, 32)), ((14, 28, 8, 8), (1, 15, 31))])
def test_ir_arranged_kernel_long_runs(shape, grid):
    """More patches than CTAs: several patches per CTA, runs that wrap around rows and images, both weight layouts; patch
    counts that do not divide by the SM count (1536 = 148 x 10.4: runs of 10 and 11 patches, none empty)."""
    Cin, hid, Cout, ps = shape
    B, fh, fw = grid
    hp = Cin * hid + 9 * hid + hid * Cout
    x = _rand((B, Cin, fh * ps, fw * ps), 80).to(DEV, torch.bfloat16)
    w = _rand((B, hp, fh, fw), 81, 0.3).to(DEV, torch.bfloat16)
    bns = [_bn(hid, 82), _bn(hid, 83), _bn(Cout, 84)]
    ref = orc.patch_ir(x.float().cpu(), w.float().cpu(), hid, Cout, *bns)
    y = _run_arranged(x, w, hid, Cout, bns)
    assert rel_err(y.float().cpu(), ref) < IR2_TOL
    y2 = _run_arranged(x, ops.weights_to_patch_major(w), hid, Cout, bns)      # run-to-run and layout independent
    assert torch.equal(y, y2)


def test_ir_arranged_refuses_what_it_cannot_run():
    from hyperseg_b200 import _lib
    x = torch.zeros(1, 20, 32, 32, device=DEV, dtype=torch.bfloat16)
    wa = torch.zeros(1, 2, 2, ops.ir_arranged_row(20, 40, 8), device=DEV, dtype=torch.bfloat16)
    z = torch.zeros(40, device=DEV)
    with pytest.raises(_lib.HsbError, match="no tensor-core instantiation"):
        ops.patch_ir_arranged(x, wa, 40, 8, z, z, z[:8])


# (signal channels, slice start, slice width, groups, head outputs, first output of the block, Cin, hid, Cout, patch)
ARRANGED_HEADS = [(1280, 0, 320, 4, 4216, 0, 34, 68, 19, 16), (1280, 0, 192, 16, 2352, 0, 24, 48, 16, 8),
                  (1280, 768, 512, 16, 3680, 868, 26, 52, 19, 16), (1280, 768, 512, 16, 3680, 0, 14, 28, 8, 8),
                  (1280, 0, 128, 8, 1896, 0, 22, 44, 12, 16), (704, 4, 192, 16, 2352, 0, 24, 48, 16, 8)]


@pytest.mark.parametrize("case", ARRANGED_HEADS)
def test_arranged_head_and_block_end_to_end(case):
    """The head that writes arranged rows (all shipped inverted-residual levels, incl. the shared unify head and a slice
    that does not start on a multiple of 8) against the float64 head followed by the re-arrangement; then the fused
    block on those rows against the oracle of the whole head + block."""
    from hyperseg_b200 import _lib
    C, si, sc, g, oc, off, cin, hid, cout, ps = case
    B, fh, fw = 2, 4, 6
    hp = cin * hid + 9 * hid + hid * cout
    s = _rand((B, C, fh, fw), 1).abs().to(DEV, torch.bfloat16)
    ws = _rand((oc, sc // g, 1, 1), 2, 0.2).to(DEV, torch.bfloat16)
    bns = [_bn(hid, 3), _bn(hid, 4), _bn(cout, 5)]
    dev = [(a.to(DEV), b.to(DEV)) for a, b in bns]
    head = ops.ArrangedHead(ws, si, sc, g, off, cin, hid, cout, dev[0][0], dev[1][0], dev[2][0])
    wa = ops.signal2weights_arranged(s, head)
    assert _lib.last_kernel() == "signal2weights_tc_kernel<arranged>"
    wref = orc.signal2weights(s.float().cpu(), ws.float().cpu(), si, sc, oc, g)[:, off:off + hp]
    wa_ref = ops.ir_arrange_weights(wref.to(DEV), cin, hid, cout, dev[0][0], dev[1][0], dev[2][0]).float().cpu()
    assert rel_err(wa.float().cpu(), wa_ref) < 1e-2          # measured <= 5.8e-3 (two bf16 roundings apart)
    x = _rand((B, cin, fh * ps, fw * ps), 6).to(DEV, torch.bfloat16)
    y = ops.patch_ir_arranged(x, wa, hid, cout, dev[0][1], dev[1][1], dev[2][1])
    yref = orc.patch_ir(x.float().cpu(), wref, hid, cout, *bns)
    assert rel_err(y.float().cpu(), yref) < 2e-2             # measured <= 1.5e-2: bf16 weights + bf16 block


def test_kernel_selection_is_observable():
    """hsb_last_kernel names what an entry point launched: the reference-order bf16 block takes the tcgen05 kernel at the
    shipped shapes and the CUDA-core kernel elsewhere (fp32, residual, other shapes); heads likewise."""
    from hyperseg_b200 import _lib
    Cin, hid, Cout, ps = 34, 68, 19, 16
    hp = Cin * hid + 9 * hid + hid * Cout
    x = _rand((1, Cin, 2 * ps, 2 * ps), 1).to(DEV)
    w = ops.weights_to_patch_major(_rand((1, hp, 2, 2), 2, 0.3).to(DEV))
    bns = [tuple(t.to(DEV) for t in _bn(hid, 3)), tuple(t.to(DEV) for t in _bn(hid, 4)), tuple(t.to(DEV) for t in _bn(Cout, 5))]
    ops.patch_ir(x.bfloat16(), w.bfloat16(), hid, Cout, *bns)
    assert _lib.last_kernel() == "patch_ir_tc_kernel"
    ops.patch_ir(x, w, hid, Cout, *bns)
    assert _lib.last_kernel() == "patch_ir_kernel"
    s = _rand((1, 1280, 4, 8), 6).to(DEV)
    ws = _rand((4216, 80, 1, 1), 7, 0.2).to(DEV)
    ops.signal2weights(s.bfloat16(), ws.bfloat16(), 0, 320, 4216, 4)
    assert _lib.last_kernel() == "signal2weights_tc_kernel<resident>", _lib.last_kernel()       # tcgen05 head, signal slice resident
    ops.signal2weights(s, ws, 0, 320, 4216, 4)
    assert _lib.last_kernel() == "signal2weights_kernel"


# (Cin, Cout, fh, fw, ph, pw, B): the 1x1 levels of the shipped configurations and shapes that exercise the masks
RING_GEOMS = [(82, 64, 16, 32, 1, 1, 2), (94, 32, 16, 32, 2, 2, 2), (44, 16, 16, 32, 4, 4, 2),        # HyperSeg-M / CamVid levels 0-2
              (130, 32, 24, 48, 1, 1, 1), (62, 16, 24, 48, 2, 2, 1), (26, 8, 12, 24, 4, 4, 1),        # HyperSeg-S Cityscapes
              (82, 64, 18, 24, 1, 1, 3), (20, 6, 5, 16, 1, 1, 2), (30, 40, 3, 8, 3, 2, 2), (18, 24, 4, 16, 2, 1, 1)]


@pytest.mark.parametrize("mma", [False, True])
@pytest.mark.parametrize("geom", RING_GEOMS)
def test_conv1x1_persistent_kernels(geom, mma, monkeypatch):
    """The persistent TMA-ring kernels (bf16, patch-major rows) against the float64 oracle, with and without the fused
    BatchNorm + ReLU: the default selection (ring kernel for one-pixel patches, one-shot kernel otherwise) and the opt-in
    mma.sync variant (HSB_CONV_MMA=1) on every 1x1 level of the shipped configurations, output-channel counts that are not a
    multiple of 16, Cin that is not a multiple of 16, odd patch widths, patch counts that do not divide by the SM count; the
    general kernel on reference-layout weights as a cross-check."""
    from hyperseg_b200 import _lib
    Cin, Cout, fh, fw, ph, pw, B = geom
    monkeypatch.setenv("HSB_CONV_MMA", "1" if mma else "0")
    want = "conv1x1_mma_kernel" if mma else ("conv1x1_ring_kernel" if ph * pw == 1 else "patch_conv1x1_kernel")
    x = _rand((B, Cin, fh * ph, fw * pw), 90).to(DEV, torch.bfloat16)
    w = _rand((B, Cin * Cout, fh, fw), 91, 0.3).to(DEV, torch.bfloat16)
    scale, shift = _bn(Cout, 92)
    wl = ops.weights_to_patch_major(w)
    for fused in (True, False):
        args = (scale.to(DEV), shift.to(DEV), "relu") if fused else (None, None, "none")
        ref = orc.patch_conv1x1(x.float().cpu(), w.float().cpu(), Cout, 1, *((scale, shift, "relu") if fused else (None, None, "none")))
        y = ops.patch_conv1x1(x, wl, Cout, 1, *args)
        assert _lib.last_kernel() == want, _lib.last_kernel()
        assert rel_err(y.float().cpu(), ref) < BF16_TOL
        y_general = ops.patch_conv1x1(x, w, Cout, 1, *args)                 # reference-layout weights: the one-shot kernel
        assert _lib.last_kernel() == "patch_conv1x1_kernel"
        assert rel_err(y.float().cpu(), y_general.float().cpu()) < 1e-2


def test_ir_module_takes_the_arranged_fast_path_in_bf16():
    """HyperPatchInvertedResidual (own head) under autocast: head -> arranged rows -> restage-free kernel, and the result
    agrees with the module's fp32 path (CUDA-core kernels)."""
    from hyperseg_b200 import _lib
    from hyperseg_b200.nn.hyperseg_v1_0 import HyperPatchInvertedResidual
    layer = HyperPatchInvertedResidual(34, 19, expand_ratio=2)
    layer.init_signal2weights(320, 0, 4)
    deterministic_init(layer, 5).eval().to(DEV)
    x, s = _rand((2, 34, 64, 96), 1).to(DEV), _rand((2, 1280, 4, 6), 2).to(DEV)
    with torch.no_grad():
        ref = layer(x, s)
        assert _lib.last_kernel() == "patch_ir_kernel"
        with torch.autocast("cuda", dtype=torch.bfloat16):
            y = layer(x, s)
        assert _lib.last_kernel() == "patch_ir2_kernel" and y.dtype == torch.bfloat16
    assert rel_err(y.float().cpu(), ref.cpu()) < 2e-2
