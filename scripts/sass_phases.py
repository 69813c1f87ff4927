"""Instruction and stall-sample share of each barrier-delimited phase of a kernel, from an ncu report's SASS page.
   python scripts/sass_phases.py gpurun_out/prof_ir.ncu-rep > profiles/r01_ir_phase_instructions.txt"""
import csv, io, subprocess, sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
head = rows[1]
col = {n: i for i, n in enumerate(head)}
segs, cur, tot_i, tot_s = [], dict(n=0, inst=0, samp=0, ops={}), 0, 0
for r in rows[2:]:
    if len(r) < len(head):
        continue
    src = r[col["Source"]].strip()
    inst, samp = int(r[col["Instructions Executed"]] or 0), int(r[col["# Samples"]] or 0)
    op = (src.split()[1] if src.startswith("@") else src.split()[0]).split(".")[0]
    cur["n"] += 1; cur["inst"] += inst; cur["samp"] += samp
    cur["ops"][op] = cur["ops"].get(op, 0) + inst
    tot_i += inst; tot_s += samp
    if "BAR.SYNC" in src:
        segs.append(cur); cur = dict(n=0, inst=0, samp=0, ops={})
segs.append(cur)
print(rows[0][1])
print(f"warp instructions executed {tot_i}, stall samples {tot_s}; segments end at successive BAR.SYNC instructions")
for i, sg in enumerate(segs):
    top = sorted(sg["ops"].items(), key=lambda kv: -kv[1])[:6]
    print(f"seg {i:2d}: {sg['n']:4d} SASS instr, {sg['inst'] / 1e6:7.2f} M executed ({100 * sg['inst'] / tot_i:4.1f} %), "
          f"{100 * sg['samp'] / max(tot_s, 1):4.1f} % of samples | " + " ".join(f"{k} {v / 1e6:.1f}M" for k, v in top))
