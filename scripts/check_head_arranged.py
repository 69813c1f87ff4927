"""GPU check of the arranged weight head (hsb_signal2weights_arranged_fwd) and of head + fused MetaBlock end to end:
arranged rows vs an fp64 head followed by the re-arrangement, block output vs the float64 oracle, timing of both heads."""
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


# (sig total, sig_index, sig_ch, groups, head out_ch, hp_offset, cin, hid, cout, ps)
CASES = [(1280, 0, 320, 4, 4216, 0, 34, 68, 19, 16), (1280, 0, 192, 16, 2352, 0, 24, 48, 16, 8),
         (1280, 768, 512, 16, 3680, 868, 26, 52, 19, 16), (1280, 768, 512, 16, 3680, 0, 14, 28, 8, 8),
         (1280, 0, 128, 8, 1896, 0, 22, 44, 12, 16)]
ok = True
for (C, si, sc, g, oc, off, cin, hid, cout, ps) in CASES:
    B, fh, fw = 2, 4, 6
    hp = cin * hid + 9 * hid + hid * cout
    s = rnd((B, C, fh, fw), 1).abs().to(DEV, torch.bfloat16)
    ws = rnd((oc, sc // g, 1, 1), 2, 0.2).to(DEV, torch.bfloat16)
    bns = [bn(hid, 3), bn(hid, 4), bn(cout, 5)]
    dev = [(a.to(DEV), b.to(DEV)) for a, b in bns]
    head = ops.ArrangedHead(ws, si, sc, g, off, cin, hid, cout, dev[0][0], dev[1][0], dev[2][0])
    wa = ops.signal2weights_arranged(s, head)
    assert _lib.last_kernel() == "signal2weights_tc_kernel<arranged>", _lib.last_kernel()
    # reference: fp64 head (oracle) on the bf16-rounded inputs, slice of this block, re-arranged with the scales
    wref = orc.signal2weights(s.float().cpu(), ws.float().cpu(), si, sc, oc, g)[:, off:off + hp]
    wa_ref = ops.ir_arrange_weights(wref.to(DEV), cin, hid, cout, dev[0][0], dev[1][0], dev[2][0]).float().cpu()
    e_head = float((wa.float().cpu() - wa_ref).abs().max() / wa_ref.abs().max())
    x = rnd((B, cin, fh * ps, fw * ps), 6).to(DEV, torch.bfloat16)
    y = ops.patch_ir_arranged(x, wa, hid, cout, dev[0][1], dev[1][1], dev[2][1])
    yref = orc.patch_ir(x.float().cpu(), wref, hid, cout, *bns)
    e_blk = float((y.float().cpu() - yref).abs().max() / yref.abs().max())
    good = e_head < 1.2e-2 and e_blk < 2e-2
    ok &= good
    print(f"case {(sc, g, oc, off, cin, hid, cout, ps)}: items {head.items} kpad_max {head.kpad_max}  head err {e_head:.2e}  block err {e_blk:.2e}  {'OK' if good else 'FAIL'}", flush=True)

# timing at the HyperSeg-M batch-8 shapes
B = 8
s = rnd((B, 1280, 16, 32), 1).abs().to(DEV, torch.bfloat16)
for name, (sc, g, oc, cin, hid, cout) in {"L4": (320, 4, 4216, 34, 68, 19), "L3": (192, 16, 2352, 24, 48, 16)}.items():
    ws = rnd((oc, sc // g, 1, 1), 2, 0.2).to(DEV, torch.bfloat16)
    one = torch.ones(max(hid, cout), device=DEV)
    head = ops.ArrangedHead(ws, 0, sc, g, 0, cin, hid, cout, one[:hid], one[:hid], one[:cout])
    hp = cin * hid + 9 * hid + hid * cout
    for label, fn in (("arranged", lambda: ops.signal2weights_arranged(s, head)), ("reference order", lambda: ops.signal2weights(s, ws, 0, sc, hp, g))):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            fn()
        e1.record()
        torch.cuda.synchronize()
        print(f"head {name} {label}: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us", flush=True)
sys.exit(0 if ok else 1)
