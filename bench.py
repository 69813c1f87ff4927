#!/usr/bin/env python
"""Benchmark of the HyperSeg decoder hot path on B200 -- prints ONE JSON line (see the driver contract).

    python bench.py [--gpus N] [--steps K] [--warmup W]                 # our arm (N>1: launched by torchrun)
    python bench.py --impl reference [--gpus N] --steps K --warmup W    # CPU reference arm (rank 0 only)

Metric (BASELINE.json): frames/sec of HyperSeg-M (EfficientNet-B1) at 1024x512, bf16, batch 8 per GPU, synthetic
frames and seeded random weights.  A "step" is one forward of one batch through the whole network: stock-PyTorch
encoder + weight mapper, and the decoder running on libhsb200's CUDA kernels.

  value     whole-job frames/s with the frames already in HBM, device-timed (CUDA events), max over ranks
  e2e       the same through SegmentationEngine.submit/collect: pinned host frames -> H2D -> forward -> argmax -> D2H
            labels, every step; the upload of step k+1 overlaps the forward of step k
  roofline  the dominant kernel (fused inverted-residual MetaBlock at decoder level 4) timed alone with CUDA events
            around a graph of 12 launches on cold inputs (3 rotating buffer sets, 437 MB > L2): algorithmic bytes /
            time against the measured HBM peak (MEASURED_PEAKS.json);
            `patch_conv` aggregates the five patch-wise kernels, `heads` the five weight heads
  cpu_baseline  the CPU port of the same forward (stock encoder + oracle decoder, fp32, all host threads) on a
            bounded sample (single frames); this is also what --impl reference times.
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

METRIC = "frames/sec HyperSeg-M 1024x512 bf16"
UNIT = "frames/s"
CONFIG = "hyperseg-m"
HEIGHT, WIDTH = 512, 1024


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=8, help="frames per GPU per step")
    ap.add_argument("--no-graph", action="store_true", help="do not capture the forward in a CUDA graph")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the CPU baseline sample")
    return ap.parse_args()


# --------------------------------------------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi sampled every 100 ms while a timed region runs (recipe: /opt/skills/guides/B200_PROFILING.md)."""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_id: str):
        self.gpu_id = gpu_id
        self.proc = None
        self.lines = []
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", self.gpu_id, f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._pump, daemon=True)
        self.thread.start()

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); smax.append(float(parts[1])); power.append(float(parts[2]))
            except ValueError:
                continue
            for n, v in zip(names, parts[3:7]):
                if v.lower() == "active":
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(smax), "power_w_max": max(power),
                "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------------------------------------------
# CPU arm: stock encoder + oracle decoder (the port of the reference's path) on single frames
# --------------------------------------------------------------------------------------------------------------
def cpu_forward_fps(steps: int, warmup: int, budget_s: float | None):
    import torch
    from hyperseg_b200.synthetic import build_model, synthetic_frames
    from oracle import hyperseg_oracle as orc

    # more than ~16 threads makes torch's CPU kernels slower on the many-core bench hosts (measured: 8 thr 29 ms,
    # 16 thr 25 ms, 32 thr 47 ms, 128 thr 9.5 s per 128x256 frame), so the baseline uses its best setting
    cores = min(os.cpu_count() or 1, 16)
    torch.set_num_threads(cores)
    model = build_model(CONFIG, seed=0)
    frames = [synthetic_frames(1, HEIGHT, WIDTH, seed=2 + i) for i in range(2)]
    times = []
    with torch.no_grad(), orc.use_oracle_ops(dtype=torch.float32):
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
    return {"fps": len(times) / total, "frames": len(times), "seconds": total, "cores": cores,
            "ms_per_frame": 1e3 * total / len(times)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    r = cpu_forward_fps(args.steps, args.warmup, None)
    sample = (f"{r['frames']} single-frame forwards (batch 1 of the batch-{args.batch} step), fp32, "
              f"torch CPU kernels + oracle decoder, {r['cores']} threads")
    line = {
        "impl": "reference", "metric": METRIC, "value": r["fps"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_frame"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "HyperSeg-M (EfficientNet-B1) Cityscapes 1024x512 inference, CPU port of the "
                               "reference path, one frame per step", "resolution": [HEIGHT, WIDTH]},
        "cpu_baseline": {"value": r["fps"], "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": sample},
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
        with open(os.path.join(REPO, "profiles", "r01_traffic.json")) as f:
            return float(json.load(f)["dram_bytes_per_launch"][tag])
    except Exception:
        return None


def kernel_rooflines(batch: int, iters: int = 12):
    """Time each decoder kernel of HyperSeg-M alone at its real shape.  Inputs rotate over 3 buffer sets so that a
    launch never finds its operands in L2 (one set of the level-4 kernel alone is 146 MB > 126 MB L2)."""
    import torch
    from hyperseg_b200 import ops
    dev = "cuda"
    dt = torch.bfloat16
    es = 2
    P = 16 * 32
    g = torch.Generator().manual_seed(0)

    def rnd(*shape, scale=1.0):
        return (torch.randn(*shape, generator=g) * scale).to(dev, dt)

    def bn(n):
        return ((torch.rand(n, generator=g) + 0.5).to(dev), (torch.randn(n, generator=g) * 0.1).to(dev))

    def time_it(fn_sets):
        """Average device time of one launch.  The launches (rotating over the cold buffer sets) are captured in a
        CUDA graph and replayed, so the CUDA events bracket device work only -- with eager launches the Python/ctypes
        call path (tens of microseconds) would dominate kernels this short."""
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
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(side):
                e0.record(side)
                graph.replay()
                e1.record(side)
            side.synchronize()
            times.append(e0.elapsed_time(e1) / iters)
        times.sort()
        return sum(times[1:-1]) / len(times[1:-1]), times[len(times) // 2]

    out = {}
    sets = 3
    # 1x1 levels: (Cin, Cout, H, W)
    for name, (cin, cout, h, w) in {"L0_conv1x1": (82, 64, 16, 32), "L1_conv1x1": (94, 32, 32, 64),
                                    "L2_conv1x1": (44, 16, 64, 128)}.items():
        fns = []
        for _ in range(sets):
            x = rnd(batch, cin, h, w)
            wt = ops.weights_to_patch_major(rnd(batch, cin * cout, 16, 32, scale=0.3))
            sc, sh = bn(cout)
            fns.append(lambda x=x, wt=wt, sc=sc, sh=sh, cout=cout: ops.patch_conv1x1(x, wt, cout, 1, sc, sh, "relu"))
        avg, med = time_it(fns)
        nbytes = es * (cin * h * w + cin * cout * P + cout * h * w) * batch
        out[name] = {"ms": avg, "ms_median": med, "bytes": nbytes}
    for name, (cin, hid, cout, h, w) in {"L3_ir": (24, 48, 16, 128, 256), "L4_ir": (34, 68, 19, 256, 512)}.items():
        hp = cin * hid + 9 * hid + hid * cout
        fns = []
        for _ in range(sets):
            x = rnd(batch, cin, h, w)
            wt = ops.weights_to_patch_major(rnd(batch, hp, 16, 32, scale=0.3))
            b1, b2, b3 = bn(hid), bn(hid), bn(cout)
            fns.append(lambda x=x, wt=wt, b1=b1, b2=b2, b3=b3, hid=hid, cout=cout: ops.patch_ir(x, wt, hid, cout, b1, b2, b3))
        avg, med = time_it(fns)
        nbytes = es * (cin * h * w + hp * P + cout * h * w) * batch
        out[name] = {"ms": avg, "ms_median": med, "bytes": nbytes}
    for name, (sc_, groups, hp) in {"L0_head": (416, 32, 5248), "L1_head": (224, 16, 3008), "L2_head": (128, 8, 704),
                                    "L3_head": (192, 16, 2352), "L4_head": (320, 4, 4216)}.items():
        fns = []
        for _ in range(sets):
            s = rnd(batch, 1280, 16, 32).abs()
            ws = rnd(hp, sc_ // groups, 1, 1, scale=0.2)
            fns.append(lambda s=s, ws=ws, sc_=sc_, hp=hp, groups=groups: ops.signal2weights(s, ws, 0, sc_, hp, groups))
        avg, med = time_it(fns)
        nbytes = es * (sc_ * P * batch + hp * sc_ // groups + hp * P * batch)
        out[name] = {"ms": avg, "ms_median": med, "bytes": nbytes}
    return out


# --------------------------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from hyperseg_b200 import _lib, ops
    from hyperseg_b200.engine import SegmentationEngine
    from hyperseg_b200.synthetic import build_model, synthetic_frames

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (our arm) needs a CUDA device: the decoder kernels have no CPU path")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    _lib.load()
    B = args.batch

    model = build_model(CONFIG, seed=0)
    engine = SegmentationEngine(model, B, HEIGHT, WIDTH, device=f"cuda:{local}", dtype=torch.bfloat16,
                                use_graph=not args.no_graph)
    rotate = 4                                  # 4 x 50 MB of frames; activations per step are > 1 GB (> L2)
    host_frames = [synthetic_frames(B, HEIGHT, WIDTH, seed=10 + rank * rotate + i).pin_memory() for i in range(rotate)]
    dev_frames = [f.to(f"cuda:{local}") for f in host_frames]
    stream = engine.stream

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def device_step(i):
        with torch.cuda.stream(stream):
            engine.frames_dev.copy_(dev_frames[i % rotate], non_blocking=True)     # D2D, frames already in HBM
        engine.step()

    props = torch.cuda.get_device_properties(local)
    gpu_id = f"GPU-{props.uuid}" if getattr(props, "uuid", None) else str(local)
    sampler = ClockSampler(gpu_id)

    # ---- value: device-timed, inputs resident ----
    for i in range(args.warmup):
        device_step(i)
    barrier()
    sampler.start()
    t_wall = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(args.steps):
        device_step(i)
    e1.record(stream)
    barrier()
    wall_ms = 1e3 * (time.perf_counter() - t_wall)
    dev_ms = e0.elapsed_time(e1)
    clocks = sampler.stop()

    # ---- e2e: host frames in, host labels out ----
    # submit()/collect(): every step uploads its own frames from pinned host memory and reads its own label map back;
    # the upload of step k+1 overlaps the forward of step k (two batches in flight)
    for i in range(min(3, args.warmup)):
        engine(host_frames[i % rotate])
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

    # ---- whole-box result (outside the timed regions): NCCL all-reduce of the confusion matrix ----
    whole_box = None
    if world > 1:
        from hyperseg_b200 import dist as hdist
        ncls = 19
        pred = engine.labels
        target = torch.roll(pred, shifts=1, dims=-1)                 # synthetic "ground truth" of the right shape
        local_mat = hdist.confusion_matrix(pred, target, ncls)
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a0.record()
        total_mat = hdist.all_reduce_confusion(local_mat.clone())
        a1.record()
        torch.cuda.synchronize()
        whole_box = {"miou": hdist.miou(total_mat)[0], "confusion_allreduce_ms": a0.elapsed_time(a1),
                     "pixels": int(total_mat.sum().item())}

    times = torch.tensor([dev_ms, e2e_ms, wall_ms], device=f"cuda:{local}", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms, wall_ms = times.tolist()
    frames_total = world * B * args.steps
    value = frames_total / (dev_ms / 1e3)
    e2e_value = frames_total / (e2e_ms / 1e3)

    line = None
    if rank == 0:
        peak, peak_src = measured_hbm_peak()
        roof = kernel_rooflines(B)
        top = roof["L4_ir"]
        achieved = top["bytes"] / (top["ms"] * 1e-3) / 1e9
        conv_keys = [k for k in roof if k.endswith("conv1x1") or k.endswith("_ir")]
        head_keys = [k for k in roof if k.endswith("_head")]

        def agg(keys):
            b = sum(roof[k]["bytes"] for k in keys)
            ms = sum(roof[k]["ms"] for k in keys)
            gbs = b / (ms * 1e-3) / 1e9
            return {"bytes": b, "ms": ms, "achieved_gbs": gbs, "frac": gbs / peak}

        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "HyperSeg-M (EfficientNet-B1) Cityscapes 1024x512 bf16 inference, batch 8 per GPU "
                                   "(BASELINE.json configs[1])",
                       "batch_per_gpu": B, "global_batch": B * world, "resolution": [HEIGHT, WIDTH],
                       "parallelism": f"dp{world}", "cuda_graph": engine.graph is not None,
                       "l2": f"{rotate} rotating input batches ({rotate * B * 3 * HEIGHT * WIDTH * 4 / 1e6:.0f} MB) and "
                             ">1 GB of activations per step exceed the 126 MB L2; no explicit flush",
                       "weights": "seeded random (hyperseg_b200.synthetic.deterministic_init)"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": B * 3 * HEIGHT * WIDTH * 4,
                    "d2h_bytes_per_step": B * HEIGHT * WIDTH, "ms_per_step": e2e_ms / args.steps,
                    "result": "uint8 argmax label map", "api": "SegmentationEngine.submit/collect, two batches in flight"},
            "gpu_launches": engine.launches_per_step * args.steps,
            "gpu_launches_per_step": engine.launches_per_step,
            "wall_ms_per_step": wall_ms / args.steps,
            "roofline": {"kernel": "hsb_patch_ir_fwd @ decoder level 4 (B x 34 x 256 x 512 -> 19 ch, 16x16 patches)",
                         "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": ncu_traffic("ir") if B == 8 else None,
                         "traffic_source": "profiles/r01_traffic.json (ncu --set full: dram read + write bytes per launch)",
                         "peak_source": peak_src, "bytes_per_launch": top["bytes"],
                         "ms_per_launch": top["ms"]},
            "patch_conv": agg(conv_keys), "heads": agg(head_keys), "whole_box": whole_box,
            "kernels": {k: {"ms": round(v["ms"], 5), "GBps": round(v["bytes"] / (v["ms"] * 1e-3) / 1e9, 1)}
                        for k, v in roof.items()},
        }
        if world == 1 and not args.no_cpu_baseline:
            r = cpu_forward_fps(64, 1, args.cpu_seconds)
            line["cpu_baseline"] = {
                "value": r["fps"], "unit": UNIT, "cores": r["cores"], "kind": "port",
                "sample": f"{r['frames']} single-frame 1024x512 forwards in {r['seconds']:.1f} s (fp32, stock torch "
                          f"encoder + oracle decoder, {r['cores']} threads of {os.cpu_count()} cores)"}
        else:
            line["cpu_baseline"] = None
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if line is not None:
        print(json.dumps(line), flush=True)
    return 0


def main():
    args = parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
