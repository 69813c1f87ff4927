"""Training path (BASELINE config 4): gradients against the reference's autograd.

tests/golden/grads.npz holds gradients produced by the unmodified reference (make_golden.py).  On CPU the
differentiable oracle is checked against them (this also pins the host-side autograd plumbing of the mirror
modules); on the GPU (-m gpu) the CUDA backward kernels, reached through torch.autograd.Function wrappers in
hyperseg_b200.ops, are checked against the same vectors -- op by op, and for one whole HyperSeg-L training step
(loss, parameter gradients of heads / weight mapper / decoder BatchNorms / encoder, BatchNorm running statistics).
"""
import os

import numpy as np
import pytest
import torch

import cases
from conftest import GOLDEN, mirror_namespace, rel_err
from hyperseg_b200.synthetic import CONFIGS, build_model, deterministic_init, synthetic_frames
from oracle import hyperseg_oracle as orc


@pytest.fixture(scope="module")
def golden_grads():
    return np.load(os.path.join(GOLDEN, "grads.npz"))


def _run_case(name, device):
    case = cases.OP_CASES[name]
    m = cases.build_op_module(mirror_namespace(), case)
    deterministic_init(m, cases.case_seed(name)).eval().to(device)
    x, w = cases.op_inputs(name, case, m.hyper_params)
    x = x.to(device).requires_grad_(True)
    w = w.to(device).requires_grad_(True)
    y = m(x, w)
    (y * cases.grad_probe(name, y.shape).to(device)).sum().backward()
    head = [p.grad for n, p in m.named_parameters() if n.endswith("signal2weights.weight")]
    return x.grad, w.grad, (head[0] if head else None)


@pytest.mark.parametrize("name", cases.GRAD_CASES)
def test_oracle_gradients_match_reference(name, golden_grads):
    with orc.use_oracle_ops():
        dx, dw, dhead = _run_case(name, "cpu")
    assert rel_err(dx, golden_grads[f"{name}/dx"]) < 5e-6
    assert rel_err(dw, golden_grads[f"{name}/dw"]) < 5e-6
    if dhead is not None:
        assert rel_err(dhead, golden_grads[f"{name}/dhead"]) < 5e-6


@pytest.mark.gpu
@pytest.mark.parametrize("name", cases.GRAD_CASES)
def test_cuda_gradients_match_reference(name, golden_grads):
    dx, dw, dhead = _run_case(name, "cuda")
    assert rel_err(dx.cpu(), golden_grads[f"{name}/dx"]) < 2e-5
    assert rel_err(dw.cpu(), golden_grads[f"{name}/dw"]) < 2e-5
    if dhead is not None:
        assert rel_err(dhead.cpu(), golden_grads[f"{name}/dhead"]) < 2e-5


def _run_train_case(name, device):
    case = cases.OP_CASES[name]
    m = cases.build_op_module(mirror_namespace(), case)
    deterministic_init(m, cases.case_seed(name)).train().to(device)
    x, w = cases.op_inputs(name, case, m.hyper_params)
    x = x.to(device).requires_grad_(True)
    w = w.to(device).requires_grad_(True)
    y = m(x, w)
    (y * cases.grad_probe(name, y.shape).to(device)).sum().backward()
    out = {"y": y.detach(), "dx": x.grad, "dw": w.grad}
    out.update({n: b for n, b in m.named_buffers() if n.endswith("running_mean") or n.endswith("running_var")})
    out.update({"d_" + n: p.grad for n, p in m.named_parameters() if p.grad is not None})
    return out


def _check_train_case(name, out, golden, tol):
    keys = [k[len(name) + 7:] for k in golden.files if k.startswith(name + "/train/")]
    assert {"y", "dx", "dw"} <= set(keys) and any(k.endswith("running_var") for k in keys)
    for k in keys:
        assert k in out, k
        assert rel_err(out[k].cpu(), golden[f"{name}/train/{k}"]) < tol, k


@pytest.mark.parametrize("name", cases.TRAIN_OP_CASES)
def test_oracle_train_mode_blocks_match_reference(name, golden_grads):
    """Blocks in train() mode: batch-statistics BatchNorm between the patch convolutions (for the v1_0 inverted
    residual block this is the stage-wise path of HyperPatchInvertedResidual._run_stagewise)."""
    with orc.use_oracle_ops():
        out = _run_train_case(name, "cpu")
    _check_train_case(name, out, golden_grads, 2e-5)


@pytest.mark.gpu
@pytest.mark.parametrize("name", cases.TRAIN_OP_CASES)
def test_cuda_train_mode_blocks_match_reference(name, golden_grads):
    out = _run_train_case(name, "cuda")
    _check_train_case(name, out, golden_grads, 1e-4)


@pytest.fixture
def cpu_drop_connect(monkeypatch):
    """Draw the drop-connect and dropout masks from the CPU generator (fp32) whatever the device, as the
    reference's CPU run did: the CUDA generator yields another stream for the same seed."""
    from hyperseg_b200.nn import efficientnet
    monkeypatch.setattr(efficientnet, "_uniform_per_sample",
                        lambda y: torch.rand([y.shape[0], 1, 1, 1]).to(device=y.device, dtype=y.dtype))
    cpu_dropout = torch.nn.functional.dropout

    def dropout(input, p=0.5, training=True, inplace=False):       # the backbone's feature dropout (train mode)
        if not training or p == 0.0:
            return input
        return input * cpu_dropout(torch.ones(input.shape), p, True).to(device=input.device, dtype=input.dtype)
    monkeypatch.setattr(torch.nn.functional, "dropout", dropout)


def _train_step(key, device, autocast=False):
    tc = cases.TRAIN_STEPS[key][0]
    cfg = CONFIGS[tc["config"]]
    model = build_model(tc["config"], seed=0).train().to(device)
    x = synthetic_frames(tc["B"], tc["H"], tc["W"]).to(device)
    labels = cases.train_labels(tc, cfg["num_classes"]).to(device)
    torch.manual_seed(tc["seed"])
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
        loss = torch.nn.functional.cross_entropy(model(x).float(), labels, ignore_index=255)
    loss.backward()
    return model, loss


def _check_train(key, model, loss, golden, tol, loss_tol=1e-5):
    # the loss agrees to ~1e-7; parameter gradients go through train-mode BatchNorm with a batch of 2 at 1x1..4x4
    # resolution, which amplifies fp32 rounding differences ~1e3-fold (a 1e-6 input perturbation moves these
    # gradients by 1e-3..7e-3): two correct fp32 implementations differ by a few 1e-3
    _, params, bn_name = cases.TRAIN_STEPS[key]
    assert abs(loss.item() - golden[f"{key}/loss"][0]) < loss_tol * abs(golden[f"{key}/loss"][0])
    named = dict(model.named_parameters())
    for pn in params:
        g = named[pn].grad
        assert g is not None, pn
        ref_norm = golden[f"{key}/{pn}/norm"][0]
        assert abs(g.double().norm().item() - ref_norm) < tol * ref_norm, pn
        ref_head = torch.from_numpy(golden[f"{key}/{pn}/head"])
        assert (g.flatten()[:64].cpu() - ref_head).abs().max().item() < tol * max(ref_head.abs().max().item(), 1e-2 * ref_norm), pn
    mean = dict(model.named_buffers())[bn_name]
    assert rel_err(mean.cpu(), golden[f"{key}/bn_mean"]) < tol


@pytest.mark.parametrize("key", list(cases.TRAIN_STEPS))
def test_oracle_training_step_matches_reference(key, golden_grads):
    with orc.use_oracle_ops():
        model, loss = _train_step(key, "cpu")
    _check_train(key, model, loss, golden_grads, 2e-2)


@pytest.mark.gpu
@pytest.mark.parametrize("key", list(cases.TRAIN_STEPS))
def test_cuda_training_step_matches_reference(key, golden_grads, cpu_drop_connect):
    """One HyperSeg-L (hyperseg_v0_1) / HyperSeg-M (hyperseg_v1_0) training step on the CUDA forward + backward
    kernels, fp32."""
    old = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        model, loss = _train_step(key, "cuda")
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
    _check_train(key, model, loss, golden_grads, 2e-2, loss_tol=1e-4)


@pytest.mark.gpu
@pytest.mark.parametrize("key", list(cases.TRAIN_STEPS))
def test_cuda_training_step_bf16_autocast_runs_and_is_close(key, golden_grads, cpu_drop_connect):
    """bf16 autocast step: finite gradients everywhere, loss within 3 %, whole-model gradient norm of the sampled
    parameters within 35 % (bf16 rounding through batch-2 BatchNorms; see _check_train)."""
    model, loss = _train_step(key, "cuda", autocast=True)
    assert abs(loss.item() - golden_grads[f"{key}/loss"][0]) < 3e-2 * golden_grads[f"{key}/loss"][0]
    named = dict(model.named_parameters())
    assert all(torch.isfinite(p.grad).all() for p in model.parameters() if p.grad is not None)
    ours = sum(named[pn].grad.double().norm().item() ** 2 for pn in cases.TRAIN_STEPS[key][1]) ** 0.5
    ref = sum(golden_grads[f"{key}/{pn}/norm"][0] ** 2 for pn in cases.TRAIN_STEPS[key][1]) ** 0.5
    assert abs(ours - ref) < 0.35 * ref


@pytest.mark.gpu
def test_fused_ir_kernel_itself_is_forward_only():
    from hyperseg_b200 import ops
    x = torch.zeros(1, 4, 8, 8, device="cuda", requires_grad=True)
    w = torch.zeros(1, 4 * 8 + 72 + 8 * 4, 1, 1, device="cuda")
    bn = (torch.ones(8, device="cuda"), torch.zeros(8, device="cuda"))
    with pytest.raises(NotImplementedError, match="forward-only"):
        ops.patch_ir(x, w, 8, 4, bn, bn, (torch.ones(4, device="cuda"), torch.zeros(4, device="cuda")))
