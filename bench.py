#!/usr/bin/env python
"""Benchmark of the HyperSeg decoder hot path on B200 -- prints ONE JSON line (see the driver contract).

    python bench.py [--gpus N] [--steps K] [--warmup W]                 # our arm (N>1: launched by torchrun)
    python bench.py --impl reference [--gpus N] --steps K --warmup W    # CPU reference arm (rank 0 only)
    python bench.py --config {m,s-city,s-camvid,l-voc-train} [--batch B] [--sweep]      # the other BASELINE configs

Default metric (BASELINE.json configs[1]): frames/sec of HyperSeg-M (EfficientNet-B1) at 1024x512, bf16, batch 8 per GPU,
synthetic frames and seeded random weights.  A "step" is one forward of one batch through the whole network: stock-PyTorch
encoder + weight mapper, and the decoder running on libhsb200's CUDA kernels.

  value     whole-job frames/s with the frames already in HBM, device-timed (CUDA events), max over ranks
  e2e       the same through SegmentationEngine.submit/collect: pinned host frames -> H2D -> forward -> argmax -> D2H
            labels, every step; the upload of step k+1 overlaps the forward of step k
  roofline  the dominant kernel (fused inverted-residual MetaBlock at decoder level 4) timed alone with CUDA events
            around a graph of 12 launches on cold inputs (3 rotating buffer sets, > L2): algorithmic bytes / time against
            the measured HBM peak (MEASURED_PEAKS.json); `kernels` has the same for every decoder kernel of the
            configuration (and every tcgen05 instantiation), `patch_conv` / `heads` aggregate them
  gpu_reference  the reference's own operator sequence on the same B200: its unmodified model (oracle/_ref, vendored by
            oracle/make_ref.py) in eager fp32 and under bf16 autocast, and its per-level ops (oracle/torch_gpu_baseline.py)
  cpu_baseline  the reference (oracle/_ref) -- or, when it has not been vendored, the CPU port (stock encoder + oracle
            decoder) -- in fp32 on the host cores, on a bounded sample (single frames); also what --impl reference times.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

UNIT = "frames/s"

# BASELINE.json configs; decoder levels per SURVEY section 3d: 1x1 levels (Cin, Cout, scale) / inverted-residual levels
# (Cin, hid, Cout, scale), scale = resolution divisor; heads (signal channels, groups, head outputs, hp) in level order
BENCH_CONFIGS = {
    "m": dict(model="hyperseg-m", res=(512, 1024), batch=8, metric="frames/sec HyperSeg-M 1024x512 bf16",
              workload="HyperSeg-M (EfficientNet-B1) Cityscapes 1024x512 bf16 inference, batch 8 per GPU (BASELINE.json configs[1])",
              conv=[(82, 64, 32), (94, 32, 16), (44, 16, 8)], ir=[(24, 48, 16, 4), (34, 68, 19, 2)],
              heads=[(416, 32, 5248, 5248), (224, 16, 3008, 3008), (128, 8, 704, 704), (192, 16, 2352, 2352), (320, 4, 4216, 4216)]),
    "s-city": dict(model="hyperseg-s-cityscapes", res=(768, 1536), batch=4, metric="frames/sec HyperSeg-S 1536x768 bf16",
                   workload="HyperSeg-S (EfficientNet-B1, unify) Cityscapes 1536x768 bf16 inference, batch 4 per GPU (BASELINE.json configs[2])",
                   conv=[(130, 32, 32), (62, 16, 16), (26, 8, 8)], ir=[(14, 28, 8, 4), (26, 52, 19, 2)],
                   heads=[(576, 32, 4160, 4160), (128, 16, 992, 992), (64, 8, 208, 208), (512, 16, 3680, 3676)]),
    "s-camvid": dict(model="hyperseg-s-camvid", res=(576, 768), batch=8, metric="frames/sec HyperSeg-S CamVid 768x576 bf16",
                     workload="HyperSeg-S (EfficientNet-B1) CamVid 768x576 bf16 inference (BASELINE.json configs[4])",
                     conv=[(82, 64, 32), (94, 32, 16), (44, 16, 8)], ir=[(24, 48, 16, 4), (22, 44, 12, 2)],
                     heads=[(448, 64, 5248, 5248), (256, 32, 3008, 3008), (256, 32, 704, 704), (192, 16, 2352, 2352), (128, 8, 1896, 1892)]),
    "l-voc-train": dict(model="hyperseg-l-voc", res=(512, 512), batch=8, metric="images/sec HyperSeg-L VOC 512x512 bf16 training step",
                        workload="HyperSeg-L (EfficientNet-B3, hyperseg_v0_1) VOC 512x512 training step (forward + backward + Adam), "
                                 "8 images per GPU, bf16 autocast, gradients all-reduced by DDP (BASELINE.json configs[3])"),
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="m", choices=sorted(BENCH_CONFIGS))
    ap.add_argument("--batch", type=int, default=None, help="frames per GPU per step (default: the configuration's)")
    ap.add_argument("--sweep", action="store_true", help="also time batch sizes 1..64 (value only) and report them in `sweep`")
    ap.add_argument("--no-graph", action="store_true", help="do not capture the forward in a CUDA graph")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-reference", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the CPU baseline sample")
    return ap.parse_args()


# --------------------------------------------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock, power and throttle reasons sampled every few milliseconds WHILE a timed region runs, through NVML in a
    thread of this process (the recipe's nvidia-smi loop needs ~100 ms per sample, longer than the 90 ms timed region, and
    its polling slowed the host-side copies of the e2e region down).  Falls back to one nvidia-smi query."""

    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))

    def __init__(self, index: int, uuid: str | None = None, period_s: float = 0.004):
        self.index, self.uuid, self.period = index, uuid, period_s
        self.rows, self.thread, self.running, self.handle, self.nvml = [], None, False, None, None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.handle = (pynvml.nvmlDeviceGetHandleByUUID(self.uuid.encode() if isinstance(self.uuid, str) else self.uuid)
                           if self.uuid else pynvml.nvmlDeviceGetHandleByIndex(self.index))
        except Exception:  # noqa: BLE001
            self.handle = None
            return
        self.running = True
        self.thread = threading.Thread(target=self._pump, daemon=True)
        self.thread.start()

    def _pump(self):
        n = self.nvml
        while self.running:
            try:
                sm = n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)
                mx = n.nvmlDeviceGetMaxClockInfo(self.handle, n.NVML_CLOCK_SM)
                pw = n.nvmlDeviceGetPowerUsage(self.handle) / 1e3
                rs = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle) if hasattr(n, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                self.rows.append((time.perf_counter(), sm, mx, pw, rs))
            except Exception:  # noqa: BLE001
                pass
            time.sleep(self.period)

    def stop(self, window=None):
        self.running = False
        if self.thread is not None:
            self.thread.join(timeout=1)
        rows = [r for r in self.rows if window is None or window[0] <= r[0] <= window[1]]
        if not rows:
            return self._fallback()
        reasons = sorted({name for _, _, _, _, rs in rows for name, bit in self.REASONS if rs & bit})
        return {"sm_mhz": statistics.median(r[1] for r in rows), "sm_max_mhz": max(r[2] for r in rows),
                "power_w_max": max(r[3] for r in rows), "samples": len(rows), "reasons": reasons, "how": "NVML, in-process thread"}

    def _fallback(self):
        try:
            out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=clocks.sm,clocks.max.sm,power.draw",
                                  "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=10).stdout.split(",")
            return {"sm_mhz": float(out[0]), "sm_max_mhz": float(out[1]), "power_w_max": float(out[2]), "samples": 1,
                    "reasons": [], "how": "one nvidia-smi query after the region (NVML unavailable)"}
        except Exception:  # noqa: BLE001
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}


# --------------------------------------------------------------------------------------------------------------
# CPU arm: the reference itself (oracle/_ref) when vendored, else the CPU port (stock encoder + oracle decoder)
# --------------------------------------------------------------------------------------------------------------
def build_reference_model(cfg_name: str):
    """The unmodified reference model of a configuration with this repo's seeded weights, or None when the reference has
    not been vendored into oracle/_ref (oracle/make_ref.py)."""
    from hyperseg_b200.synthetic import CONFIGS, deterministic_init
    from oracle import make_ref
    if not make_ref.available():
        return None
    cfg = CONFIGS[cfg_name]
    mod = make_ref.load_reference(cfg["module"])
    kwargs = {k: (list(v) if isinstance(v, list) else v) for k, v in cfg["kwargs"].items()}
    model = mod.hyperseg_efficientnet(cfg["model_name"], pretrained=False, num_classes=cfg["num_classes"], **kwargs)
    return deterministic_init(model, 0).eval()


def cpu_forward_fps(bc, steps: int, warmup: int, budget_s: float | None):
    import torch
    from hyperseg_b200.synthetic import build_model, synthetic_frames
    from oracle import hyperseg_oracle as orc
    # more than ~16 threads makes torch's CPU kernels slower on the many-core bench hosts (measured: 8 thr 29 ms,
    # 16 thr 25 ms, 32 thr 47 ms, 128 thr 9.5 s per 128x256 frame), so the baseline uses its best setting
    cores = min(os.cpu_count() or 1, 16)
    torch.set_num_threads(cores)
    H, W = bc["res"]
    ref = build_reference_model(bc["model"])
    kind = "reference" if ref is not None else "port"
    model = ref if ref is not None else build_model(bc["model"], seed=0)
    frames = [synthetic_frames(1, H, W, seed=2 + i) for i in range(2)]
    times = []
    import contextlib
    ctx = contextlib.nullcontext() if ref is not None else orc.use_oracle_ops(dtype=torch.float32)
    with torch.no_grad(), ctx:
        for i in range(warmup):
            model(frames[i % 2])
        t_begin = time.perf_counter()
        for i in range(steps):
            t0 = time.perf_counter()
            model(frames[i % 2])
            times.append(time.perf_counter() - t0)
            if budget_s is not None and time.perf_counter() - t_begin > budget_s and len(times) >= 3:
                break
    total = sum(times)
    return {"fps": len(times) / total, "frames": len(times), "seconds": total, "cores": cores, "kind": kind,
            "ms_per_frame": 1e3 * total / len(times)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    bc = BENCH_CONFIGS[args.config if args.config != "l-voc-train" else "m"]
    H, W = bc["res"]
    r = cpu_forward_fps(bc, args.steps, args.warmup, None)
    what = ("the unmodified reference model (oracle/_ref), torch CPU kernels" if r["kind"] == "reference"
            else "CPU port: stock torch encoder + oracle decoder")
    sample = (f"{r['frames']} single-frame {W}x{H} forwards (batch 1 of the batch-{bc['batch']} step), fp32, {what}, "
              f"{r['cores']} threads")
    line = {
        "impl": "reference", "metric": bc["metric"], "value": r["fps"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_frame"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": bc["workload"] + " -- reference arm: CPU forward, one frame per step", "resolution": [H, W]},
        "cpu_baseline": {"value": r["fps"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"], "sample": sample},
        "e2e": {"value": r["fps"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# --------------------------------------------------------------------------------------------------------------
# per-kernel roofline (device events, cold inputs)
# --------------------------------------------------------------------------------------------------------------
def measured_hbm_peak():
    path = os.path.join(REPO, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(tag: str):
    """DRAM bytes per launch of a kernel from the committed ncu --set full capture (profiles/), or None."""
    try:
        with open(os.path.join(REPO, "profiles", "r02_traffic.json")) as f:
            return float(json.load(f)["dram_bytes_per_launch"][tag])
    except Exception:
        return None


def time_graph(fn_sets, iters=12, reps=7):
    """Average device time (ms) of one launch.  The launches (rotating over cold buffer sets) are captured in a CUDA graph
    and replayed, so the CUDA events -- recorded on the stream the graph runs on -- bracket device work only; with eager
    launches the Python/ctypes call path (tens of microseconds) would dominate kernels this short."""
    import torch
    for f in fn_sets:                       # warm-up (sets func attributes / loads modules / packs head weights)
        f()
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        for f in fn_sets:
            f()
        side.synchronize()
        with torch.cuda.graph(graph, stream=side):
            for i in range(iters):
                fn_sets[i % len(fn_sets)]()
    torch.cuda.synchronize()
    times = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(side):
            e0.record(side)
            graph.replay()
            e1.record(side)
        side.synchronize()
        times.append(e0.elapsed_time(e1) / iters)
    times.sort()
    return sum(times[1:-1]) / len(times[1:-1])


def time_eager(fn, reps=5):
    """Device time (ms) of an eager torch call sequence (the GPU baseline: large ATen kernels, launch overhead is theirs)."""
    import torch
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def kernel_rooflines(bc, batch: int, with_gpu_reference: bool):
    """Time each decoder kernel of the configuration alone at its real shape.  Inputs rotate over 3 buffer sets so that a
    launch never finds its operands in L2.  Algorithmic bytes per SURVEY section 8(d)."""
    import torch
    from hyperseg_b200 import ops
    dev, dt, es = "cuda", torch.bfloat16, 2
    H, W = bc["res"]
    fh, fw = H // 32, W // 32
    P = fh * fw
    g = torch.Generator().manual_seed(0)

    def rnd(*shape, scale=1.0):
        return (torch.randn(*shape, generator=g) * scale).to(dev, dt)

    def bn(n):
        return ((torch.rand(n, generator=g) + 0.5).to(dev), (torch.randn(n, generator=g) * 0.1).to(dev))

    out, gpu_ref = {}, {}
    sets = 3
    if with_gpu_reference:
        from oracle import torch_gpu_baseline as tgb
    for li, (cin, cout, div) in enumerate(bc["conv"]):
        h, w = H // div, W // div
        fns, keep = [], None
        for _ in range(sets):
            x = rnd(batch, cin, h, w)
            wt = ops.weights_to_patch_major(rnd(batch, cin * cout, fh, fw, scale=0.3))
            sc, sh = bn(cout)
            fns.append(lambda x=x, wt=wt, sc=sc, sh=sh, cout=cout: ops.patch_conv1x1(x, wt, cout, 1, sc, sh, "relu"))
            keep = (x, wt, sc, sh)
        ms = time_graph(fns)
        from hyperseg_b200 import _lib as _l
        out[f"L{li}_conv1x1"] = {"ms": ms, "bytes": es * (cin * h * w + cin * cout * P + cout * h * w) * batch, "kernel": _l.last_kernel()}
        if with_gpu_reference:
            x, wt, sc, sh = keep
            b = tgb.make_bn(cout, sc, sh, dev)
            gpu_ref[f"L{li}_conv1x1"] = {
                "bf16_ms": time_eager(lambda: tgb.patch_conv1x1(x, wt, cout, b)),
                "fp32_ms": time_eager(lambda: tgb.patch_conv1x1(x.float(), wt.float(), cout, b))}
    nconv = len(bc["conv"])
    for li, (cin, hid, cout, div) in enumerate(bc["ir"]):
        h, w = H // div, W // div
        hp = cin * hid + 9 * hid + hid * cout
        fns_new, fns_raw, keep = [], [], None
        for _ in range(sets):
            x = rnd(batch, cin, h, w)
            wt = ops.weights_to_patch_major(rnd(batch, hp, fh, fw, scale=0.3))
            b1, b2, b3 = bn(hid), bn(hid), bn(cout)
            wa = ops.ir_arrange_weights(wt, cin, hid, cout, b1[0], b2[0], b3[0])
            fns_new.append(lambda x=x, wa=wa, b1=b1, b2=b2, b3=b3, hid=hid, cout=cout: ops.patch_ir_arranged(x, wa, hid, cout, b1[1], b2[1], b3[1]))
            fns_raw.append(lambda x=x, wt=wt, b1=b1, b2=b2, b3=b3, hid=hid, cout=cout: ops.patch_ir(x, wt, hid, cout, b1, b2, b3))
            keep = (x, wt, b1, b2, b3)
        nbytes = es * (cin * h * w + hp * P + cout * h * w) * batch
        out[f"L{nconv + li}_ir"] = {"ms": time_graph(fns_new), "bytes": nbytes, "kernel": "patch_ir2_kernel (arranged weight rows)"}
        out[f"L{nconv + li}_ir_reference_order"] = {"ms": time_graph(fns_raw), "bytes": nbytes, "kernel": "patch_ir_tc_kernel (round 1)"}
        if with_gpu_reference:
            x, wt, b1, b2, b3 = keep
            m1, m2, m3 = tgb.make_bn(hid, *b1, dev), tgb.make_bn(hid, *b2, dev), tgb.make_bn(cout, *b3, dev)
            gpu_ref[f"L{nconv + li}_ir"] = {
                "bf16_ms": time_eager(lambda: tgb.patch_ir(x, wt, hid, cout, m1, m2, m3), reps=3),
                "fp32_ms": time_eager(lambda: tgb.patch_ir(x.float(), wt.float(), hid, cout, m1.float(), m2.float(), m3.float()), reps=3)}
    ir_levels = {nconv + li: v for li, v in enumerate(bc["ir"])}
    offset = 0
    for li, (sc_, groups, och, hp) in enumerate(bc["heads"]):
        fns, keep = [], None
        for _ in range(sets):
            s = rnd(batch, 1280, fh, fw).abs()
            ws = rnd(och, sc_ // groups, 1, 1, scale=0.2)
            fns.append(lambda s=s, ws=ws, sc_=sc_, hp=hp, groups=groups: ops.signal2weights(s, ws, 0, sc_, hp, groups))
            keep = (s, ws)
        nbytes = es * (sc_ * P * batch + hp * sc_ // groups + hp * P * batch)
        from hyperseg_b200 import _lib as _l
        out[f"L{li}_head"] = {"ms": time_graph(fns), "bytes": nbytes, "kernel": _l.last_kernel() + " (reference-order rows)"}
        # the arranged variant feeds the inverted-residual levels (a shared unify head feeds several: one pack per level)
        feeds = [lv for lv in ir_levels if lv == li] if len(bc["heads"]) == nconv + len(bc["ir"]) else (list(ir_levels) if li == len(bc["heads"]) - 1 else [])
        off = 0
        for lv in feeds:
            cin, hid, cout, _ = ir_levels[lv]
            hp_l = cin * hid + 9 * hid + hid * cout
            s, ws = keep
            one = torch.ones(max(hid, cout), device=dev)
            try:
                heads = [ops.ArrangedHead(ws, 0, sc_, groups, off, cin, hid, cout, one[:hid], one[:hid], one[:cout]) for _ in range(sets)]
            except Exception:
                break
            ss = [rnd(batch, 1280, fh, fw).abs() for _ in range(sets)]
            fns = [lambda s=ss[i], hd=heads[i]: ops.signal2weights_arranged(s, hd) for i in range(sets)]
            out[f"L{lv}_head_arranged"] = {"ms": time_graph(fns), "bytes": es * (sc_ * P * batch + hp_l * sc_ // groups + hp_l * P * batch),
                                           "kernel": "signal2weights_tc_kernel<arranged>"}
            off += hp_l
        if with_gpu_reference:
            s, ws = keep
            gpu_ref[f"L{li}_head"] = {"bf16_ms": time_eager(lambda: tgb.head(s, ws, 0, sc_, hp, groups)),
                                      "fp32_ms": time_eager(lambda: tgb.head(s.float(), ws.float(), 0, sc_, hp, groups))}
    return out, gpu_ref


def gpu_reference_model(bc, batch: int):
    """The unmodified reference model on this GPU, timed the way hyperseg/test_fps.py:172-191 does (synchronize, forward,
    synchronize) but with BatchNorm kept (test_fps strips it, which changes the outputs): eager fp32 and bf16 autocast."""
    import torch
    from hyperseg_b200.synthetic import synthetic_frames
    model = build_reference_model(bc["model"])
    if model is None:
        return None
    H, W = bc["res"]
    model = model.cuda()
    torch.backends.cudnn.benchmark = True
    frames = synthetic_frames(batch, H, W, seed=3).cuda()
    res = {}
    with torch.no_grad():
        for name, ctx in (("eager_fp32", None), ("autocast_bf16", torch.autocast("cuda", dtype=torch.bfloat16))):
            def fwd():
                if ctx is None:
                    return model(frames)
                with ctx:
                    return model(frames)
            try:
                ms = time_eager(fwd, reps=5)
                res[name] = {"ms_per_step": ms, "fps": batch / (ms * 1e-3)}
            except Exception as e:  # noqa: BLE001  (e.g. out of memory on a shared box)
                res[name] = {"error": str(e)[:120]}
    del model
    torch.cuda.empty_cache()
    return res


# --------------------------------------------------------------------------------------------------------------
# our arm: inference configurations
# --------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from hyperseg_b200 import _lib, ops
    from hyperseg_b200.engine import SegmentationEngine
    from hyperseg_b200.synthetic import build_model, synthetic_frames

    bc = BENCH_CONFIGS[args.config]
    HEIGHT, WIDTH = bc["res"]
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (our arm) needs a CUDA device: the decoder kernels have no CPU path")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    _lib.load()
    B = args.batch or bc["batch"]

    model = build_model(bc["model"], seed=0)
    engine = SegmentationEngine(model, B, HEIGHT, WIDTH, device=f"cuda:{local}", dtype=torch.bfloat16,
                                use_graph=not args.no_graph)
    rotate = 4                                  # rotating input batches; activations per step are > 1 GB (> L2)
    host_frames = [synthetic_frames(B, HEIGHT, WIDTH, seed=10 + rank * rotate + i).pin_memory() for i in range(rotate)]
    dev_frames = [f.to(f"cuda:{local}") for f in host_frames]
    stream = engine.stream

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def device_step(eng, frames, i):
        with torch.cuda.stream(eng.stream):
            eng.frames_dev.copy_(frames[i % rotate], non_blocking=True)     # D2D, frames already in HBM
        eng.step()

    props = torch.cuda.get_device_properties(local)
    sampler = ClockSampler(local, f"GPU-{props.uuid}" if getattr(props, "uuid", None) else None)

    # ---- value: device-timed, inputs resident ----
    for i in range(args.warmup):
        device_step(engine, dev_frames, i)
    barrier()
    sampler.start()                                     # clocks are sampled during the device-timed region only
    t_wall = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(args.steps):
        device_step(engine, dev_frames, i)
    e1.record(stream)
    barrier()
    t_wall_end = time.perf_counter()
    wall_ms = 1e3 * (t_wall_end - t_wall)
    dev_ms = e0.elapsed_time(e1)
    clocks = sampler.stop(window=(t_wall, t_wall_end))

    # ---- e2e: host frames in, host labels out ----
    # submit()/collect(): every step uploads its own frames from pinned host memory and reads its own label map back;
    # the upload of step k+1 overlaps the forward of step k (two batches in flight)
    for i in range(max(3, min(5, args.warmup))):     # warm-up through the same calls (creates the staging / pinned buffers)
        engine.submit(host_frames[i % rotate])
        engine.collect()
    barrier()
    e2e_start = time.perf_counter()
    checksum = 0
    for i in range(args.steps):
        engine.submit(host_frames[i % rotate])
        if i > 0:
            checksum += int(engine.collect()[0, 0, 0])
    checksum += int(engine.collect()[0, 0, 0])
    barrier()
    e2e_ms = 1e3 * (time.perf_counter() - e2e_start)

    # ---- the same loop with frames as they come off a camera / decoder: 8-bit RGB, normalised on the device ----
    engine8 = SegmentationEngine(model, B, HEIGHT, WIDTH, device=f"cuda:{local}", dtype=torch.bfloat16,
                                 use_graph=not args.no_graph, input_dtype=torch.uint8)
    mean = torch.tensor([0.485, 0.456, 0.406]).view(1, 3, 1, 1)
    std = torch.tensor([0.229, 0.224, 0.225]).view(1, 3, 1, 1)
    host_u8 = [((f * std + mean) * 255).round().clamp(0, 255).to(torch.uint8).pin_memory() for f in host_frames]
    for i in range(max(3, min(5, args.warmup))):
        engine8.submit(host_u8[i % rotate])
        engine8.collect()
    barrier()
    u8_start = time.perf_counter()
    for i in range(args.steps):
        engine8.submit(host_u8[i % rotate])
        if i > 0:
            checksum += int(engine8.collect()[0, 0, 0])
    checksum += int(engine8.collect()[0, 0, 0])
    barrier()
    u8_ms = 1e3 * (time.perf_counter() - u8_start)
    del engine8

    # ---- whole-box result (outside the timed regions): NCCL collectives over NVLink ----
    whole_box = None
    if world > 1:
        from hyperseg_b200 import dist as hdist
        ncls = engine.logits.shape[1]
        pred = engine.labels
        target = torch.roll(pred, shifts=1, dims=-1)                 # synthetic "ground truth" of the right shape
        local_mat = hdist.confusion_matrix(pred, target, ncls)
        a0, a1, a2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        torch.cuda.synchronize()
        a0.record()
        total_mat = hdist.all_reduce_confusion(local_mat.clone())
        a1.record()
        gathered = hdist.gather_logits(engine.logits, total_items=world * B)      # north_star's all-gather of the logits
        a2.record()
        torch.cuda.synchronize()
        whole_box = {"miou": hdist.miou(total_mat)[0], "confusion_allreduce_ms": a0.elapsed_time(a1),
                     "logits_allgather_ms": a1.elapsed_time(a2), "logits_allgather_bytes": gathered.numel() * gathered.element_size(),
                     "logits_allgather_shape": list(gathered.shape), "pixels": int(total_mat.sum().item())}
        del gathered

    times = torch.tensor([dev_ms, e2e_ms, wall_ms, u8_ms], device=f"cuda:{local}", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms, wall_ms, u8_ms = times.tolist()
    frames_total = world * B * args.steps
    value = frames_total / (dev_ms / 1e3)
    e2e_value = frames_total / (e2e_ms / 1e3)

    # ---- optional batch sweep (BASELINE.json configs[4]): device-timed value at other batch sizes ----
    sweep = None
    if args.sweep:
        sweep = {}
        for b in (1, 2, 4, 8, 16, 32, 64):
            try:
                eng = SegmentationEngine(model, b, HEIGHT, WIDTH, device=f"cuda:{local}", dtype=torch.bfloat16, use_graph=not args.no_graph)
                fr = [synthetic_frames(b, HEIGHT, WIDTH, seed=50 + i).to(f"cuda:{local}") for i in range(rotate)]
                for i in range(3):
                    device_step(eng, fr, i)
                barrier()
                s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s0.record(eng.stream)
                n = max(5, min(args.steps, 200 // b))
                for i in range(n):
                    device_step(eng, fr, i)
                s1.record(eng.stream)
                barrier()
                t = torch.tensor([s0.elapsed_time(s1)], device=f"cuda:{local}", dtype=torch.float64)
                if world > 1:
                    dist.all_reduce(t, op=dist.ReduceOp.MAX)
                sweep[str(b)] = {"fps": world * b * n / (t.item() * 1e-3), "ms_per_step": t.item() / n}
                del eng, fr
                torch.cuda.empty_cache()
            except Exception as e:  # noqa: BLE001
                sweep[str(b)] = {"error": str(e)[:100]}

    line = None
    if rank == 0:
        peak, peak_src = measured_hbm_peak()
        with_ref = world == 1 and not args.no_gpu_reference
        roof, gpu_ops = kernel_rooflines(bc, B, with_ref)
        top_key = f"L{len(bc['conv']) + len(bc['ir']) - 1}_ir"
        top = roof[top_key]
        achieved = top["bytes"] / (top["ms"] * 1e-3) / 1e9
        conv_keys = [k for k in roof if k.endswith("conv1x1") or k.endswith("_ir")]
        # heads as the model runs them: arranged rows for the inverted-residual levels, reference-order rows elsewhere
        head_keys = [k for k in roof if k.endswith("_head") and k.replace("_head", "_head_arranged") not in roof] + \
                    [k for k in roof if k.endswith("_head_arranged")]

        def agg(keys):
            b = sum(roof[k]["bytes"] for k in keys)
            ms = sum(roof[k]["ms"] for k in keys)
            gbs = b / (ms * 1e-3) / 1e9
            return {"bytes": b, "ms": ms, "achieved_gbs": gbs, "frac": gbs / peak}

        cin, hid, cout, div = bc["ir"][-1]
        line = {
            "metric": bc["metric"], "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": bc["workload"], "batch_per_gpu": B, "global_batch": B * world, "resolution": [HEIGHT, WIDTH],
                       "parallelism": f"dp{world}", "cuda_graph": engine.graph is not None,
                       "l2": f"{rotate} rotating input batches ({rotate * B * 3 * HEIGHT * WIDTH * 4 / 1e6:.0f} MB) and "
                             ">1 GB of activations per step exceed the 126 MB L2; no explicit flush",
                       "weights": "seeded random (hyperseg_b200.synthetic.deterministic_init)"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": B * 3 * HEIGHT * WIDTH * 4,
                    "d2h_bytes_per_step": B * HEIGHT * WIDTH, "ms_per_step": e2e_ms / args.steps,
                    "result": "uint8 argmax label map", "api": "SegmentationEngine.submit/collect, two batches in flight",
                    "frames": "float32, normalised on the host (the reference's input convention)"},
            "e2e_uint8_frames": {"value": frames_total / (u8_ms / 1e3), "unit": UNIT, "h2d_bytes_per_step": B * 3 * HEIGHT * WIDTH,
                                 "d2h_bytes_per_step": B * HEIGHT * WIDTH, "ms_per_step": u8_ms / args.steps,
                                 "frames": "uint8 RGB, (x / 255 - mean) / std on the device (SegmentationEngine(input_dtype=torch.uint8))"},
            "gpu_launches": engine.launches_per_step * args.steps,
            "gpu_launches_per_step": engine.launches_per_step,
            "wall_ms_per_step": wall_ms / args.steps,
            "roofline": {"kernel": f"hsb_patch_ir_arranged_fwd (patch_ir2_kernel) @ last decoder level (B x {cin} x {HEIGHT // div} x "
                                   f"{WIDTH // div} -> {cout} ch, hidden {hid}, 16x16 patches)",
                         "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": ncu_traffic("ir2_l4") if (B == 8 and args.config == "m") else None,
                         "traffic_source": "profiles/r02_traffic.json (ncu --set full: dram read + write bytes per launch)",
                         "peak_source": peak_src, "bytes_per_launch": top["bytes"], "ms_per_launch": top["ms"]},
            "patch_conv": agg(conv_keys), "heads": agg(head_keys), "whole_box": whole_box,
            "kernels": {k: {"ms": round(v["ms"], 5), "GBps": round(v["bytes"] / (v["ms"] * 1e-3) / 1e9, 1),
                            "frac": round(v["bytes"] / (v["ms"] * 1e-3) / 1e9 / peak, 4), **({"kernel": v["kernel"]} if "kernel" in v else {})}
                        for k, v in roof.items()},
        }
        if sweep is not None:
            line["sweep"] = sweep
        if with_ref:
            # other tcgen05 instantiations (BASELINE configs 3 and 5) at their own full sizes: every shape gets a fraction
            extra = {}
            for name in ("s-city", "s-camvid"):
                if name == args.config:
                    continue
                try:
                    r2, _ = kernel_rooflines(dict(BENCH_CONFIGS[name], conv=[], heads=[]), BENCH_CONFIGS[name]["batch"], False)
                    for k, v in r2.items():
                        if k.endswith("_ir"):
                            lv = len(BENCH_CONFIGS[name]["conv"]) + int(k[1])
                            extra[f"{name}_L{lv}_ir"] = {"ms": round(v["ms"], 5), "GBps": round(v["bytes"] / (v["ms"] * 1e-3) / 1e9, 1),
                                                         "frac": round(v["bytes"] / (v["ms"] * 1e-3) / 1e9 / peak, 4)}
                except Exception as e:  # noqa: BLE001
                    extra[name] = {"error": str(e)[:100]}
            line["kernels_other_configs"] = extra
            ref_model = gpu_reference_model(bc, B)
            per_op = {}
            for k, v in gpu_ops.items():
                ours_key = k if k in roof else None
                if k.endswith("_head") and k.replace("_head", "_head_arranged") in roof:
                    ours_key = k.replace("_head", "_head_arranged")
                ours = roof[ours_key]["ms"] if ours_key else None
                per_op[k] = {"reference_bf16_ms": round(v["bf16_ms"], 4), "reference_fp32_ms": round(v["fp32_ms"], 4),
                             "ours_ms": round(ours, 5) if ours else None,
                             "speedup_vs_best_reference_mode": round(min(v["bf16_ms"], v["fp32_ms"]) / ours, 1) if ours else None}
            line["gpu_reference"] = {
                "what": "the reference's operator sequence (pad / unfold / F.conv2d(groups = B*P) / BatchNorm / ReLU6 / permute) on this GPU: "
                        "per level from oracle/torch_gpu_baseline.py, whole model = unmodified reference (oracle/_ref) timed like test_fps.py",
                "ops": per_op, "model": ref_model,
                "model_speedup_vs_best_reference_mode": (round(value / max(v["fps"] for v in ref_model.values() if "fps" in v), 2)
                                                         if ref_model and any("fps" in v for v in ref_model.values()) else None)}
        if world == 1 and not args.no_cpu_baseline:
            r = cpu_forward_fps(bc, 64, 1, args.cpu_seconds)
            what = "unmodified reference model (oracle/_ref)" if r["kind"] == "reference" else "stock torch encoder + oracle decoder"
            line["cpu_baseline"] = {
                "value": r["fps"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"],
                "sample": f"{r['frames']} single-frame {WIDTH}x{HEIGHT} forwards in {r['seconds']:.1f} s (fp32, {what}, "
                          f"{r['cores']} threads of {os.cpu_count()} cores)"}
        else:
            line["cpu_baseline"] = None
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if line is not None:
        print(json.dumps(line), flush=True)
    return 0


# --------------------------------------------------------------------------------------------------------------
# our arm: HyperSeg-L training step (BASELINE.json configs[3])
# --------------------------------------------------------------------------------------------------------------
def run_train(args):
    import torch
    import torch.distributed as dist
    from hyperseg_b200 import _lib, ops
    from hyperseg_b200.synthetic import build_model, synthetic_frames

    bc = BENCH_CONFIGS["l-voc-train"]
    H, W = bc["res"]
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    _lib.load()
    B = args.batch or bc["batch"]
    dev = f"cuda:{local}"
    model = build_model(bc["model"], seed=0).to(dev).train()
    net = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local]) if world > 1 else model
    # the reference's training defaults: CrossEntropyLoss(ignore_index=255), Adam lr 1e-4 betas (0.5, 0.999)
    # (hyperseg/train.py:30,66, configs/train/vocsbd_efficientnet_b3_hyperseg-l.py)
    opt = torch.optim.Adam(model.parameters(), lr=1e-4, betas=(0.5, 0.999))
    crit = torch.nn.CrossEntropyLoss(ignore_index=255)
    g = torch.Generator().manual_seed(7 + rank)
    frames = [synthetic_frames(B, H, W, seed=20 + rank * 4 + i).to(dev) for i in range(4)]
    labels = [torch.randint(0, 21, (B, H, W), generator=g).to(dev) for _ in range(4)]
    host_frames = [f.cpu().pin_memory() for f in frames]
    host_labels = [l.cpu().pin_memory() for l in labels]

    def step(x, y):
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            out = net(x)
        if out.shape[-2:] != y.shape[-2:]:
            out = torch.nn.functional.interpolate(out, y.shape[-2:], mode="bilinear", align_corners=False)
        loss = crit(out.float(), y)
        loss.backward()
        opt.step()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    props = torch.cuda.get_device_properties(local)
    sampler = ClockSampler(local, f"GPU-{props.uuid}" if getattr(props, "uuid", None) else None)
    # DDP rebuilds its gradient buckets after the first iteration and the caching allocator keeps growing for a few more:
    # at least five untimed steps (a 4-GPU run with three still had a 136 ms step inside the timed region, 107 ms after it)
    args.warmup = max(args.warmup, 5)
    for i in range(args.warmup):
        step(frames[i % 4], labels[i % 4])
    barrier()
    before = ops.launch_count()
    sampler.start()
    t0w = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        loss = step(frames[i % 4], labels[i % 4])
    e1.record()
    barrier()
    dev_ms = e0.elapsed_time(e1)
    clocks = sampler.stop(window=(t0w, time.perf_counter()))
    launches = ops.launch_count() - before
    # e2e: inputs and labels from pinned host memory every step, the loss read back every step
    barrier()
    t0 = time.perf_counter()
    last = 0.0
    for i in range(args.steps):
        x = host_frames[i % 4].to(dev, non_blocking=True)
        y = host_labels[i % 4].to(dev, non_blocking=True)
        last = float(step(x, y).item())
    barrier()
    t1 = time.perf_counter()
    e2e_ms = 1e3 * (t1 - t0)
    times = torch.tensor([dev_ms, e2e_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms = times.tolist()
    if rank == 0:
        n = world * B * args.steps
        line = {"metric": bc["metric"], "value": n / (dev_ms * 1e-3), "unit": "images/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": {"workload": bc["workload"], "batch_per_gpu": B, "global_batch": B * world, "resolution": [H, W],
                           "parallelism": f"ddp{world}", "optimizer": "Adam lr 1e-4 betas (0.5, 0.999)", "loss": "CrossEntropyLoss(ignore_index=255)",
                           "l2": "4 rotating input batches; activations per step exceed the 126 MB L2"},
                "clocks": clocks,
                "e2e": {"value": n / (e2e_ms * 1e-3), "unit": "images/s", "h2d_bytes_per_step": B * 3 * H * W * 4 + B * H * W * 8,
                        "d2h_bytes_per_step": 4, "ms_per_step": e2e_ms / args.steps, "result": "loss (float)", "last_loss": last},
                "gpu_launches": launches, "gpu_launches_per_step": launches // max(args.steps, 1),
                "roofline": None, "cpu_baseline": None}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    args = parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.config == "l-voc-train":
        return run_train(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
