"""Summarise gpurun_out/prof_<tag>.ncu-rep (ncu --set full, one launch each; scripts/gpu_final.sh) into
profiles/<round>_ncu_full_summary.txt and profiles/<round>_traffic.json.   python scripts/summarize_ncu.py r01"""
import csv, io, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TAGS = [("ir2", "hsb_patch_ir_arranged_fwd (patch_ir2_kernel), HyperSeg-M level 4 (B=8, 34->68->19, 256x512, 16x16 patches, bf16)"),
        ("ir2_3", "hsb_patch_ir_arranged_fwd (patch_ir2_kernel), level 3 (B=8, 24->48->16, 128x256, 8x8 patches)"),
        ("head4a", "hsb_signal2weights_arranged_fwd, level-4 head writing arranged rows (320 -> 4704-element rows, groups 4, B=8)"),
        ("head3a", "hsb_signal2weights_arranged_fwd, level-3 head writing arranged rows (192 -> 2352, groups 16)"),
        ("ir", "hsb_patch_ir_fwd, reference-order tcgen05 path (round 1), HyperSeg-M level 4 (B=8, 34->68->19, 256x512, 16x16 patches, bf16)"),
        ("ir3", "hsb_patch_ir_fwd, tcgen05 path, level 3 (B=8, 24->48->16, 128x256, 8x8 patches)"),
        ("head4", "hsb_signal2weights_packed_fwd, level-4 head (320 -> 4216, groups 4, B=8)"),
        ("head0", "hsb_signal2weights_packed_fwd, level-0 head (416 -> 5248, groups 32)"),
        ("conv0", "hsb_patch_conv1x1_fwd, level 0 (82 -> 64, 1x1 patches; round 2: conv1x1_ring_kernel)"),
        ("conv1", "hsb_patch_conv1x1_fwd, level 1 (94 -> 32, 2x2 patches)"),
        ("conv2", "hsb_patch_conv1x1_fwd, level 2 (44 -> 16, 4x4 patches)"),
        ("epi", "hsb_bias_act_nhwc_fwd, encoder epilogue (shift + swish + SE partial sums), B=8, 96 ch, 256x512, bf16")]
METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
           "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
           "launch__occupancy_limit_registers", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
           "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
           "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"]
TO_BYTES = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main(rnd):
    out, traffic = [], {}
    for tag, title in TAGS:
        rep = os.path.join(ROOT, "gpurun_out", f"prof_{tag}.ncu-rep")
        if not os.path.exists(rep):
            continue
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        head, units, vals = rows[0], rows[1], rows[2]
        col = {h: i for i, h in enumerate(head)}
        out.append(f"== {title}\n   kernel: {vals[col['Kernel Name']]}")
        rd = wr = 0.0
        for m in METRICS:
            if m in col:
                out.append(f"   {m:86s} {vals[col[m]]} {units[col[m]]}")
                if m.startswith("dram__bytes"):
                    b = float(vals[col[m]].replace(",", "")) * TO_BYTES.get(units[col[m]], 1.0)
                    rd, wr = (b, wr) if "read" in m else (rd, b)
        traffic[tag] = rd + wr
        if tag == "ir2":
            traffic["ir2_l4"] = rd + wr          # the key bench.py reads for roofline.traffic
        out.append("")
    hdr = (f"ncu --set full --clock-control none --import-source on, one launch each at the HyperSeg-M batch-8 shape "
           f"(scripts/run_kernel.py, scripts/gpu_final.sh, summarised by scripts/summarize_ncu.py); B200, round {rnd[1:]}.\n"
           f"Times under ncu are cold-cache single launches; bench.py's CUDA-event numbers are the ones reported.\n")
    with open(os.path.join(ROOT, "profiles", f"{rnd}_ncu_full_summary.txt"), "w") as f:
        f.write(hdr + "\n" + "\n".join(out))
    with open(os.path.join(ROOT, "profiles", f"{rnd}_traffic.json"), "w") as f:
        json.dump({"source": f"profiles/{rnd}_ncu_full_summary.txt (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per launch)",
                   "dram_bytes_per_launch": traffic}, f, indent=1)
    print("\n".join(out[:30]))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "r01")
