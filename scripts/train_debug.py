import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import numpy as np, torch, cases
from hyperseg_b200.synthetic import CONFIGS, build_model, synthetic_frames
from hyperseg_b200.nn import efficientnet
from oracle.hyperseg_oracle import use_oracle_ops
import contextlib
efficientnet._uniform_per_sample = lambda y: torch.rand([y.shape[0], 1, 1, 1]).to(device=y.device, dtype=y.dtype)
_cpu_dropout = torch.nn.functional.dropout
def _dropout(input, p=0.5, training=True, inplace=False):
    if not training or p == 0.0: return input
    return input * _cpu_dropout(torch.ones(input.shape), p, True).to(device=input.device, dtype=input.dtype)
torch.nn.functional.dropout = _dropout
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
tc = cases.TRAIN_CASE; cfg = CONFIGS[tc["config"]]
def run(oracle):
    dev = "cpu" if oracle else "cuda"
    model = build_model(tc["config"], 0).train().to(dev)
    x = synthetic_frames(tc["B"], tc["H"], tc["W"]).to(dev); labels = cases.train_labels(tc, cfg["num_classes"]).to(dev)
    outs = {}
    def hook(name):
        def f(m, i, o):
            if isinstance(o, torch.Tensor): outs[name] = o.detach().double().cpu().clone()
        return f
    for n, m in model.named_modules(): m.register_forward_hook(hook(n))
    torch.manual_seed(tc["seed"])
    with (use_oracle_ops() if oracle else contextlib.nullcontext()):
        loss = torch.nn.functional.cross_entropy(model(x), labels, ignore_index=255)
        loss.backward()
    return loss.item(), outs, {n: p.grad.double().cpu() for n, p in model.named_parameters() if p.grad is not None}
l0, o0, g0 = run(True); l1, o1, g1 = run(False)
print("loss oracle-on-gpu", l0, "cuda", l1)
k = 0
for n in o0:
    if n in o1 and o0[n].shape == o1[n].shape:
        e = ((o0[n] - o1[n]).norm() / (o0[n].norm() + 1e-30)).item()
        if e > 1e-4 and k < 25: print(f"fwd {n:50s} rel {e:.2e}"); k += 1
k = 0
for n in g0:
    e = ((g0[n] - g1[n]).norm() / (g0[n].norm() + 1e-30)).item()
    if e > 1e-3 and k < 40: print(f"grad {n:50s} rel {e:.2e} norm {g0[n].norm().item():.3e}"); k += 1
