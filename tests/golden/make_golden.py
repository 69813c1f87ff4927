"""Generate the golden parity vectors by running the UNMODIFIED reference (needs /root/reference).

    python tests/golden/make_golden.py

The reference repository holds no golden vectors or known-answer tests of its own, so
the oracle and the CUDA path are pinned against outputs of the reference's Python modules, imported here
from /root/reference (build container only -- the GPU box has no copy) and executed on CPU in float32 on
seeded inputs.  Written files (all committed):

    tests/golden/ops.npz      op-level: inputs + reference outputs of every case in cases.OP_CASES / HEAD_CASES
    tests/golden/models.npz   whole-model logits (spatial stride 2) of cases.MODEL_CASES, weights from
                              hyperseg_b200.synthetic.deterministic_init
    tests/golden/divide.npz   divide_feature / divide_feature_legacy results and per-config head geometry
    tests/golden/grads.npz    reference autograd gradients of the op cases in cases.GRAD_CASES and of one
                              HyperSeg-L (hyperseg_v0_1) and one HyperSeg-M (hyperseg_v1_0) training step (loss, parameter
                              gradients, BN statistics); blocks in train() mode
"""
import importlib
import os
import sys
import types
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
sys.path.insert(0, HERE)
REFERENCE = "/root/reference"


def import_reference():
    if not os.path.isdir(REFERENCE):
        raise SystemExit("make_golden.py needs the reference checkout at /root/reference")
    sys.modules.setdefault("ffmpeg", types.ModuleType("ffmpeg"))    # hyperseg/utils/utils.py:9 imports it, unused here
    if REFERENCE not in sys.path:
        sys.path.insert(0, REFERENCE)
    warnings.filterwarnings("ignore", message="torch.meshgrid")
    v10 = importlib.import_module("hyperseg.models.hyperseg_v1_0")
    v01 = importlib.import_module("hyperseg.models.hyperseg_v0_1")
    mp = importlib.import_module("hyperseg.models.layers.meta_patch")
    mc = importlib.import_module("hyperseg.models.layers.meta_conv")
    return {
        "HyperPatchNoPadding": v10.HyperPatchNoPadding,
        "make_hyper_patch_conv2d_block": v10.make_hyper_patch_conv2d_block,
        "HyperPatchInvertedResidual": v10.HyperPatchInvertedResidual,
        "HyperPatchConv2d": v10.HyperPatchConv2d,
        "MetaPatchConv2d": mp.MetaPatchConv2d,
        "make_meta_patch_conv2d_block": mp.make_meta_patch_conv2d_block,
        "MetaConv2d": mc.MetaConv2d,
        "V01HyperPatchInvertedResidual": v01.HyperPatchInvertedResidual,
    }


def main():
    import cases
    from hyperseg_b200.synthetic import CONFIGS, deterministic_init, synthetic_frames

    torch.manual_seed(0)
    torch.set_grad_enabled(False)
    ns = import_reference()

    ops = {}
    for name, case in cases.OP_CASES.items():
        m = cases.build_op_module(ns, case)
        deterministic_init(m, cases.case_seed(name)).eval()
        x, w = cases.op_inputs(name, case, m.hyper_params)
        y = m(x, w)
        ops[f"{name}/x"], ops[f"{name}/w"], ops[f"{name}/y"] = x.numpy(), w.numpy(), y.numpy()
        print(f"{name:22s} x{tuple(x.shape)} w{tuple(w.shape)} -> y{tuple(y.shape)} |y|max={y.abs().max():.3f}")

    v10 = importlib.import_module("hyperseg.models.hyperseg_v1_0")
    for name, c in cases.HEAD_CASES.items():
        layer = v10.HyperPatchNoPadding(c["hp"], 1, 1)          # hyper_params == c["hp"]
        assert int(layer.hyper_params) == c["hp"]
        layer.init_signal2weights(c["sc"], c["idx"], c["groups"])
        deterministic_init(layer, cases.case_seed(name))
        s = cases.head_inputs(name, c)
        wgt = layer.apply_signal2weights(s)
        ops[f"{name}/s"], ops[f"{name}/y"] = s.numpy(), wgt.numpy()
        print(f"{name:22s} s{tuple(s.shape)} -> w{tuple(wgt.shape)} conv_out={layer.signal2weights.out_channels}")
    np.savez_compressed(os.path.join(HERE, "ops.npz"), **ops)

    # ---- gradients of the same modules (reference autograd), training path ----
    torch.set_grad_enabled(True)
    grads = {}
    for name in cases.GRAD_CASES:
        case = cases.OP_CASES[name]
        m = cases.build_op_module(ns, case)
        deterministic_init(m, cases.case_seed(name)).eval()
        x, w = cases.op_inputs(name, case, m.hyper_params)
        x.requires_grad_(True); w.requires_grad_(True)
        y = m(x, w)
        (y * cases.grad_probe(name, y.shape)).sum().backward()
        grads[f"{name}/dx"], grads[f"{name}/dw"] = x.grad.numpy(), w.grad.numpy()
        for pn, pp in m.named_parameters():
            if pn.endswith("signal2weights.weight"):
                grads[f"{name}/dhead"] = pp.grad.numpy()
        print(f"grad {name:20s} |dx|={x.grad.abs().max():.3f} |dw|={w.grad.abs().max():.3f}")

    for name in cases.TRAIN_OP_CASES:
        case = cases.OP_CASES[name]
        m = cases.build_op_module(ns, case)
        deterministic_init(m, cases.case_seed(name)).train()
        x, w = cases.op_inputs(name, case, m.hyper_params)
        x.requires_grad_(True); w.requires_grad_(True)
        y = m(x, w)
        (y * cases.grad_probe(name, y.shape)).sum().backward()
        grads[f"{name}/train/y"] = y.detach().numpy()
        grads[f"{name}/train/dx"], grads[f"{name}/train/dw"] = x.grad.numpy(), w.grad.numpy()
        for bn, bb in m.named_buffers():
            if bn.endswith("running_mean") or bn.endswith("running_var"):
                grads[f"{name}/train/{bn}"] = bb.numpy().copy()
        for pn, pp in m.named_parameters():
            if pp.grad is not None:
                grads[f"{name}/train/d_{pn}"] = pp.grad.numpy()
        print(f"train-mode {name:14s} |y|={y.abs().max():.3f} |dx|={x.grad.abs().max():.3f} |dw|={w.grad.abs().max():.3f}")

    for key, (tc, params, bn_name) in cases.TRAIN_STEPS.items():
        cfg = CONFIGS[tc["config"]]
        mod = importlib.import_module("hyperseg.models." + cfg["module"])
        kwargs = {k: (list(v) if isinstance(v, list) else v) for k, v in cfg["kwargs"].items()}
        model = mod.hyperseg_efficientnet(cfg["model_name"], pretrained=False, num_classes=cfg["num_classes"], **kwargs)
        deterministic_init(model, 0).train()
        x = synthetic_frames(tc["B"], tc["H"], tc["W"])
        labels = cases.train_labels(tc, cfg["num_classes"])
        torch.manual_seed(tc["seed"])                     # drop-connect / dropout masks
        loss = torch.nn.functional.cross_entropy(model(x), labels, ignore_index=255)
        loss.backward()
        grads[f"{key}/loss"] = np.array([loss.item()], dtype=np.float64)
        named = dict(model.named_parameters())
        for pn in params:
            g = named[pn].grad
            grads[f"{key}/{pn}/norm"] = np.array([g.double().norm().item()])
            grads[f"{key}/{pn}/head"] = g.flatten()[:64].numpy()
        grads[f"{key}/bn_mean"] = dict(model.named_buffers())[bn_name].numpy()
        print(f"{key} step ({tc['config']}) loss={loss.item():.6f} grad norms " + ' '.join(f"{named[pn].grad.norm().item():.2e}" for pn in params))
    np.savez_compressed(os.path.join(HERE, "grads.npz"), **grads)
    torch.set_grad_enabled(False)

    models = {}
    geometry = {}
    for name, mc_ in cases.MODEL_CASES.items():
        cfg = CONFIGS[mc_["config"]]
        mod = importlib.import_module("hyperseg.models." + cfg["module"])
        kwargs = {k: (list(v) if isinstance(v, list) else v) for k, v in cfg["kwargs"].items()}
        model = mod.hyperseg_efficientnet(cfg["model_name"], pretrained=False, num_classes=cfg["num_classes"], **kwargs)
        deterministic_init(model, 0).eval()
        x = synthetic_frames(mc_["B"], mc_["H"], mc_["W"])
        y = model(x)
        st = cases.MODEL_STRIDE
        models[f"{name}/logits"] = y[:, :, ::st, ::st].numpy()
        models[f"{name}/stats"] = np.array([y.mean().item(), y.std().item(), y.abs().max().item(),
                                            x.double().sum().item()], dtype=np.float64)
        models[f"{name}/argmax"] = y.argmax(1).to(torch.uint8).numpy()
        print(f"{name:24s} logits{tuple(y.shape)} std={y.std():.4f} max={y.abs().max():.4f}")
        # head geometry of the configuration (signal_index / channels / groups / out channels), depth-first
        heads = []
        for mname, m in model.named_modules():
            conv = getattr(m, "signal2weights", None)
            if conv is not None:
                hp = getattr(m, "hyper_params", getattr(m, "target_params", -1))
                heads.append((int(m.signal_index), int(m.signal_channels), int(conv.groups), int(conv.out_channels), int(hp)))
        if hasattr(model.weight_mapper, "out_conv"):
            oc = model.weight_mapper.out_conv
            for i in range(len(oc.out_channels)):
                conv = getattr(oc, f"conv_{i}")
                heads.append((int(oc._ranges[i]), int(conv.in_channels), int(conv.groups), int(conv.out_channels), -1))
        geometry[mc_["config"]] = np.array(heads, dtype=np.int64)
        keys = sorted(model.state_dict().keys())
        geometry[mc_["config"] + "/keys"] = np.array(keys)
        geometry[mc_["config"] + "/shapes"] = np.array([str(tuple(model.state_dict()[k].shape)) for k in keys])
    np.savez_compressed(os.path.join(HERE, "models.npz"), **models)

    div = {}
    v01 = importlib.import_module("hyperseg.models.hyperseg_v0_1")
    for i, (inf, outs, unit) in enumerate(cases.DIVIDE_CASES):
        div[f"v1_0/{i}"] = np.asarray(v10.divide_feature(inf, list(outs), min_unit=unit), dtype=np.int64)
        try:
            div[f"legacy/{i}"] = np.asarray(v01.divide_feature_legacy(inf, list(outs), unit), dtype=np.int64)
        except Exception as e:       # the legacy variant has inputs it cannot handle
            div[f"legacy/{i}"] = np.array([-1], dtype=np.int64)
    div.update({"geometry/" + k: v for k, v in geometry.items()})
    np.savez_compressed(os.path.join(HERE, "divide.npz"), **div)
    for f in ("ops.npz", "models.npz", "divide.npz", "grads.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)) // 1024, "KiB")


if __name__ == "__main__":
    main()
