"""GPU check of the restage-free MetaBlock kernel (hsb_patch_ir_arranged_fwd) against the float64 oracle and the
round-1 kernel, every shipped shape and border case, then A/B timing at the HyperSeg-M batch-8 shapes.

    python scripts/check_ir2.py [--time-only] [--no-time]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hyperseg_b200 import _lib, ops  # noqa: E402
from oracle import hyperseg_oracle as orc  # noqa: E402

DEV = "cuda"


def rnd(shape, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g) * scale


def bn(n, seed):
    g = torch.Generator().manual_seed(seed)
    return (torch.rand(n, generator=g) + 0.5), (torch.randn(n, generator=g) * 0.1)


def rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-12))


def run_new(x, w, hid, cout, bns):
    cin = x.shape[1]
    wa = ops.ir_arrange_weights(w, cin, hid, cout, bns[0][0].to(DEV), bns[1][0].to(DEV), bns[2][0].to(DEV))
    y = ops.patch_ir_arranged(x, wa, hid, cout, bns[0][1].to(DEV), bns[1][1].to(DEV), bns[2][1].to(DEV))
    assert _lib.last_kernel() == "patch_ir2_kernel", _lib.last_kernel()
    return y


def parity():
    worst = 0.0
    shapes = [(34, 68, 19, 16), (26, 52, 19, 16), (22, 44, 12, 16), (24, 48, 16, 8), (14, 28, 8, 8)]
    grids = [(2, 3, 5), (1, 1, 1), (3, 1, 4), (1, 2, 1), (2, 16, 24)]
    for shape in shapes:
        cin, hid, cout, ps = shape
        for B, fh, fw in grids:
            hp = cin * hid + 9 * hid + hid * cout
            x = rnd((B, cin, fh * ps, fw * ps), 70).to(DEV, torch.bfloat16)
            w = rnd((B, hp, fh, fw), 71, 0.3).to(DEV, torch.bfloat16)
            bns = [bn(hid, 72), bn(hid, 73), bn(cout, 74)]
            ref = orc.patch_ir(x.float().cpu(), w.float().cpu(), hid, cout, *bns)
            y = run_new(x, w, hid, cout, bns)
            torch.cuda.synchronize()
            y_old = ops.patch_ir(x, ops.weights_to_patch_major(w), hid, cout, *[(a.to(DEV), b.to(DEV)) for a, b in bns])
            e_new, e_old = rel(y.float().cpu(), ref), rel(y_old.float().cpu(), ref)
            worst = max(worst, e_new)
            flag = "OK" if e_new < 1.5e-2 else "FAIL"
            print(f"shape {shape} grid {(B, fh, fw)}: new {e_new:.2e}  old {e_old:.2e}  {flag}", flush=True)
            if e_new >= 1.5e-2:
                d = (y.float().cpu() - ref).abs()
                idx = torch.nonzero(d > 0.05 * ref.abs().max())
                print("   first bad:", idx[:8].tolist(), "count", idx.shape[0])
    print("worst rel err", worst)
    return worst < 1.5e-2


def sm_clock():
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(0)
        return pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
    except Exception as e:  # noqa: BLE001
        return (-1, -1)


def timing():
    B = 8
    # spin the GPU up first: short bursts after idle run at a low SM clock
    a = torch.randn(8192, 8192, device=DEV, dtype=torch.bfloat16)
    for _ in range(30):
        a @ a
    torch.cuda.synchronize()
    print("SM clock after spin-up (MHz, max):", sm_clock(), flush=True)
    cases = {"ir": (34, 68, 19, 256, 512), "ir3": (24, 48, 16, 128, 256)}
    for name, (cin, hid, cout, h, w_) in cases.items():
        fns_new, fns_old = [], []
        for k in range(3):
            x = rnd((B, cin, h, w_), k).to(DEV, torch.bfloat16)
            wt = ops.weights_to_patch_major(rnd((B, cin * hid + 9 * hid + hid * cout, 16, 32), 10 + k, 0.3).to(DEV, torch.bfloat16))
            bns = [bn(hid, 1), bn(hid, 2), bn(cout, 3)]
            dev = [(a.to(DEV), b.to(DEV)) for a, b in bns]
            wa = ops.ir_arrange_weights(wt, cin, hid, cout, dev[0][0], dev[1][0], dev[2][0])
            fns_new.append(lambda x=x, wa=wa, dev=dev: ops.patch_ir_arranged(x, wa, hid, cout, dev[0][1], dev[1][1], dev[2][1]))
            fns_old.append(lambda x=x, wt=wt, dev=dev: ops.patch_ir(x, wt, hid, cout, *dev))
        for label, fns in (("new", fns_new), ("old", fns_old)):
            with torch.no_grad():
                for f in fns:
                    f()
                torch.cuda.synchronize()
                side = torch.cuda.Stream()
                graph = torch.cuda.CUDAGraph()
                iters = 12
                with torch.cuda.stream(side):
                    for f in fns:
                        f()
                    side.synchronize()
                    with torch.cuda.graph(graph, stream=side):
                        for i in range(iters):
                            fns[i % 3]()
                torch.cuda.synchronize()
                ts = []
                for _ in range(20):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    with torch.cuda.stream(side):
                        e0.record(side)
                        graph.replay()
                        e1.record(side)
                    side.synchronize()
                    ts.append(e0.elapsed_time(e1) / iters)
                us = sorted(ts)[len(ts) // 2] * 1e3
                clk = sm_clock()
                hp = cin * hid + 9 * hid + hid * cout
                nbytes = 2 * B * (cin * h * w_ + hp * 512 + cout * h * w_)
                print(f"{name} {label}: {us:.1f} us/launch  {nbytes / us * 1e-3:.0f} GB/s  frac {nbytes / us * 1e-3 / 6539.5:.3f}  (min {min(ts) * 1e3:.1f} us, SM clock {clk[0]} MHz)", flush=True)


def graph_time(fns, iters=12, reps=15):
    with torch.no_grad():
        for f in fns:
            f()
        torch.cuda.synchronize()
        side = torch.cuda.Stream()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.stream(side):
            for f in fns:
                f()
            side.synchronize()
            with torch.cuda.graph(graph, stream=side):
                for i in range(iters):
                    fns[i % len(fns)]()
        torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(side):
                e0.record(side)
                graph.replay()
                e1.record(side)
            side.synchronize()
            ts.append(e0.elapsed_time(e1) / iters)
    return sorted(ts)[len(ts) // 2] * 1e3


def scan():
    """Launch time against patches per CTA (batch 1..8 of the 16x32 grid): least-squares fixed cost + cost per patch."""
    a = torch.randn(8192, 8192, device=DEV, dtype=torch.bfloat16)
    for _ in range(30):
        a @ a
    torch.cuda.synchronize()
    for name, (cin, hid, cout, h, w_) in {"ir": (34, 68, 19, 256, 512), "ir3": (24, 48, 16, 128, 256)}.items():
        pts = []
        for B in (1, 2, 3, 4, 6, 8):
            fns = []
            for k in range(3):
                x = rnd((B, cin, h, w_), k).to(DEV, torch.bfloat16)
                wt = ops.weights_to_patch_major(rnd((B, cin * hid + 9 * hid + hid * cout, 16, 32), 10 + k, 0.3).to(DEV, torch.bfloat16))
                dev = [(a_.to(DEV), b_.to(DEV)) for a_, b_ in [bn(hid, 1), bn(hid, 2), bn(cout, 3)]]
                wa = ops.ir_arrange_weights(wt, cin, hid, cout, dev[0][0], dev[1][0], dev[2][0])
                fns.append(lambda x=x, wa=wa, dev=dev: ops.patch_ir_arranged(x, wa, hid, cout, dev[0][1], dev[1][1], dev[2][1]))
            us = graph_time(fns)
            per_cta = -(-B * 512 // 148)
            pts.append((per_cta, us))
            print(f"{name} B={B}: {per_cta} patches per CTA, {us:.1f} us", flush=True)
        n = len(pts); sx = sum(p for p, _ in pts); sy = sum(u for _, u in pts)
        sxx = sum(p * p for p, _ in pts); sxy = sum(p * u for p, u in pts)
        b = (n * sxy - sx * sy) / (n * sxx - sx * sx); a0 = (sy - b * sx) / n
        print(f"{name}: fixed {a0:.1f} us + {b:.2f} us per patch per CTA ({b * 1965:.0f} cycles at 1965 MHz)", flush=True)


if __name__ == "__main__":
    if "--scan" in sys.argv:
        scan()
        sys.exit(0)
    ok = True
    if "--time-only" not in sys.argv:
        ok = parity()
    if "--no-time" not in sys.argv:
        timing()
    sys.exit(0 if ok else 1)
