"""The oracle (and the host-side module logic around it) against the reference's own outputs.

Every case in tests/golden/ops.npz was produced by the unmodified reference (tests/golden/make_golden.py).
Here the same module is rebuilt from hyperseg_b200.nn (identical state_dict keys -> identical seeded
parameters), executed on CPU with hyperseg_b200.ops routed to oracle/hyperseg_oracle.py, and compared.
This pins the oracle; the CUDA kernels are then compared with the oracle / the same vectors under -m gpu.
"""
import pytest
import torch

import cases
from conftest import mirror_namespace, rel_err
from hyperseg_b200.synthetic import deterministic_init
from oracle import hyperseg_oracle as orc

TOL = 2e-6      # float64 oracle vs float32 reference, relative to max |y|


@pytest.mark.parametrize("name", sorted(cases.OP_CASES))
def test_op_case(name, golden_ops):
    case = cases.OP_CASES[name]
    m = cases.build_op_module(mirror_namespace(), case)
    deterministic_init(m, cases.case_seed(name)).eval()
    x = torch.from_numpy(golden_ops[f"{name}/x"])
    w = torch.from_numpy(golden_ops[f"{name}/w"])
    # the stored inputs are the seeded ones
    xs, ws = cases.op_inputs(name, case, m.hyper_params)
    assert torch.equal(xs, x) and torch.equal(ws, w)
    with torch.no_grad(), orc.use_oracle_ops():
        y = m(x, w)
    ref = torch.from_numpy(golden_ops[f"{name}/y"])
    assert y.shape == ref.shape
    assert rel_err(y, ref) < TOL


@pytest.mark.parametrize("name", sorted(cases.HEAD_CASES))
def test_head_case(name, golden_ops):
    from hyperseg_b200.nn.hyperseg_v1_0 import HyperPatchNoPadding
    c = cases.HEAD_CASES[name]
    layer = HyperPatchNoPadding(c["hp"], 1, 1)
    layer.init_signal2weights(c["sc"], c["idx"], c["groups"])
    deterministic_init(layer, cases.case_seed(name))
    s = torch.from_numpy(golden_ops[f"{name}/s"])
    assert torch.equal(s, cases.head_inputs(name, c))
    with torch.no_grad(), orc.use_oracle_ops():
        w = layer.apply_signal2weights(s)
    ref = torch.from_numpy(golden_ops[f"{name}/y"])
    assert w.shape == ref.shape
    assert rel_err(w, ref) < TOL


def test_raw_functions_match_modules(golden_ops):
    """The oracle functions called directly (no nn.Module in between) on the weight-fed cases."""
    x = torch.from_numpy(golden_ops["nopad_groups/x"])
    w = torch.from_numpy(golden_ops["nopad_groups/w"])
    y = orc.patch_conv1x1(x, w, 12, groups=4)
    assert rel_err(y, golden_ops["nopad_groups/y"]) < TOL
    x = torch.from_numpy(golden_ops["metapatch_dil/x"])
    w = torch.from_numpy(golden_ops["metapatch_dil/w"])
    y = orc.patch_conv(x, w, 5, (3, 3), (2, 2), (2, 2), 1, "replicate")
    assert rel_err(y, golden_ops["metapatch_dil/y"]) < TOL
    x = torch.from_numpy(golden_ops["metaconv_valid/x"])
    w = torch.from_numpy(golden_ops["metaconv_valid/w"])
    y = orc.meta_conv2d(x, w, 2, (1, 3))
    assert rel_err(y, golden_ops["metaconv_valid/y"]) < TOL


def test_patch_conv_k1_equals_conv1x1():
    """the generic path with k=1 is the 1x1 path."""
    g = torch.Generator().manual_seed(5)
    x = torch.randn(2, 6, 8, 12, generator=g)
    w = torch.randn(2, 6 * 4, 2, 3, generator=g)
    a = orc.patch_conv1x1(x, w, 4)
    b = orc.patch_conv(x, w, 4, (1, 1), (0, 0))
    assert torch.allclose(a, b, atol=1e-6)


def test_ir_per_patch_loop():
    """A literal per-patch loop against the vectorised oracle, incl. the residual."""
    g = torch.Generator().manual_seed(11)
    B, Cin, hid, Cout, fh, fw, ph, pw = 1, 4, 8, 4, 2, 2, 3, 5
    H, W = fh * ph, fw * pw
    x = torch.randn(B, Cin, H, W, generator=g, dtype=torch.float64)
    hp = Cin * hid + 9 * hid + hid * Cout
    w = torch.randn(B, hp, fh, fw, generator=g, dtype=torch.float64) * 0.3
    bn = [(torch.rand(n, generator=g, dtype=torch.float64) + 0.5, torch.randn(n, generator=g, dtype=torch.float64) * 0.1)
          for n in (hid, hid, Cout)]
    y = orc.patch_ir(x, w, hid, Cout, *bn, residual=True)
    xp = torch.nn.functional.pad(x, (1, 1, 1, 1), mode="reflect")
    out = torch.zeros(B, Cout, H, W, dtype=torch.float64)
    for i in range(fh):
        for j in range(fw):
            wv = w[0, :, i, j]
            W1 = wv[:Cin * hid].view(hid, Cin)
            W2 = wv[Cin * hid:Cin * hid + 9 * hid].view(hid, 3, 3)
            W3 = wv[Cin * hid + 9 * hid:].view(Cout, hid)
            tile = xp[0, :, i * ph:i * ph + ph + 2, j * pw:j * pw + pw + 2]
            h = torch.einsum('oc,crq->orq', W1, tile)
            h = (h * bn[0][0][:, None, None] + bn[0][1][:, None, None]).clamp(0, 6)
            d = torch.zeros(hid, ph, pw, dtype=torch.float64)
            for u in range(ph):
                for v in range(pw):
                    d[:, u, v] = (W2 * h[:, u:u + 3, v:v + 3]).sum((1, 2))
            d = (d * bn[1][0][:, None, None] + bn[1][1][:, None, None]).clamp(0, 6)
            o = torch.einsum('oc,cuv->ouv', W3, d) * bn[2][0][:, None, None] + bn[2][1][:, None, None]
            out[0, :, i * ph:(i + 1) * ph, j * pw:(j + 1) * pw] = o
    out = out + x
    assert torch.allclose(y, out, atol=1e-12)


def test_gpu_baseline_operator_sequence_matches_the_oracle():
    """oracle/torch_gpu_baseline.py (the stock-ATen operator sequence bench.py times as ``gpu_reference``) computes the
    same three functions as the float64 oracle: head, 1x1 block, inverted-residual block.  fp32 on CPU, tolerance 1e-5."""
    from oracle import torch_gpu_baseline as tgb
    g = torch.Generator().manual_seed(23)
    B, fh, fw = 2, 2, 3
    # head: groups 4, truncated to hp rows
    s = torch.randn(B, 48, fh, fw, generator=g)
    ws = torch.randn(72, 4, 1, 1, generator=g)
    a = tgb.head(s, ws, 16, 16, 70, 4)
    b = orc.signal2weights(s, ws, 16, 16, 70, 4)
    assert rel_err(a, b) < 1e-5
    # 1x1 block with BatchNorm + ReLU
    x = torch.randn(B, 6, fh * 4, fw * 8, generator=g)
    w = torch.randn(B, 6 * 5, fh, fw, generator=g)
    scale, shift = torch.rand(5, generator=g) + 0.5, torch.randn(5, generator=g) * 0.1
    a = tgb.patch_conv1x1(x, w, 5, tgb.make_bn(5, scale, shift, "cpu"), relu=True)
    b = orc.patch_conv1x1(x, w, 5, scale=scale, shift=shift, act="relu")
    assert rel_err(a, b) < 1e-5
    # inverted-residual block
    cin, hid, cout = 6, 12, 5
    w = torch.randn(B, cin * hid + 9 * hid + hid * cout, fh, fw, generator=g) * 0.3
    bns = [(torch.rand(n, generator=g) + 0.5, torch.randn(n, generator=g) * 0.1) for n in (hid, hid, cout)]
    mods = [tgb.make_bn(n, sc, sh, "cpu") for n, (sc, sh) in zip((hid, hid, cout), bns)]
    with torch.no_grad():
        a = tgb.patch_ir(x, w, hid, cout, *mods)
    b = orc.patch_ir(x, w, hid, cout, *bns)
    assert rel_err(a, b) < 1e-5
