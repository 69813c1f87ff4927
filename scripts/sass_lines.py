"""Per-source-line instruction and stall-sample counts of one kernel: joins the SASS page of an ncu report (per
instruction counters, in program order) with nvdisasm's line table for the same function of the object file.

    python scripts/sass_lines.py gpurun_out/prof.ncu-rep hyperseg_b200/csrc/build/patch_ir2.o 'IR2ILi34' [units]

`units` divides the counts (e.g. 4096 patches per launch).  The object must be the build that was profiled."""
import csv
import io
import re
import subprocess
import sys
import tempfile
from collections import defaultdict

rep, obj, pattern = sys.argv[1], sys.argv[2], sys.argv[3]
units = float(sys.argv[4]) if len(sys.argv) > 4 else 1.0

raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
head = rows[1]
col = {n: i for i, n in enumerate(head)}
prof = []
for r in rows[2:]:
    if len(r) < len(head):
        continue
    stalls = {k[6:]: int(r[i] or 0) for k, i in col.items() if k.startswith("stall_") and "Not Issued" not in k}
    wf = (int(r[col["L1 Wavefronts Shared"]] or 0), int(r[col["L1 Wavefronts Shared Ideal"]] or 0))
    prof.append((r[col["Source"]].strip(), int(r[col["Instructions Executed"]] or 0), int(r[col["# Samples"]] or 0), stalls, wf))

with tempfile.TemporaryDirectory() as tmp:
    import os; subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
    import glob
    cubin = glob.glob(tmp + "/*.cubin")[0]
    dis = subprocess.run(["nvdisasm", "-g", cubin], capture_output=True, text=True).stdout.splitlines()

# locate the function's text section
start = None
for i, ln in enumerate(dis):
    if ln.startswith(".text.") and pattern in ln:
        start = i
        break
assert start is not None, "function not found"
lines = []   # (line number, opcode text) in program order
cur = 0
for ln in dis[start + 1:]:
    if ln.startswith(".text.") or ln.lstrip().startswith(".section"):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = int(m.group(2)) if m.group(1).endswith(".cu") and "patch" in m.group(1) or "signal" in m.group(1) else -int(m.group(2))
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(.*?);", ln)
    if m:
        lines.append((cur, m.group(1).strip()))
if len(lines) != len(prof):
    print(f"warning: {len(lines)} disassembled instructions vs {len(prof)} profiled", file=sys.stderr)
per = defaultdict(lambda: [0, 0, defaultdict(int), defaultdict(int), [0, 0]])
for (ln, op), (src, inst, samp, stalls, wf) in zip(lines, prof):
    per[ln][0] += inst
    per[ln][1] += samp
    for k, v in stalls.items():
        per[ln][3][k] += v
    per[ln][4][0] += wf[0]
    per[ln][4][1] += wf[1]
    o = (src.split()[1] if src.startswith("@") else src.split()[0]).split(".")[0]
    per[ln][2][o] += inst
tot_i = sum(v[0] for v in per.values())
tot_s = sum(v[1] for v in per.values())
print(f"total warp instructions {tot_i} ({tot_i / units:.0f} per unit), samples {tot_s}")
src_file = None
for ln, v in sorted(per.items(), key=lambda kv: -kv[1][0])[:70]:
    ops = " ".join(f"{k} {c / units:.0f}" for k, c in sorted(v[2].items(), key=lambda kv: -kv[1])[:5])
    print(f"line {ln:5d}: {v[0] / units:8.1f} instr/unit ({100 * v[0] / tot_i:4.1f} %)  samples {100 * v[1] / max(tot_s, 1):4.1f} %  | {ops}")

print("\n--- lines by stall samples ---")
for ln, v in sorted(per.items(), key=lambda kv: -kv[1][1])[:40]:
    st = " ".join(f"{k} {c}" for k, c in sorted(v[3].items(), key=lambda kv: -kv[1])[:4] if c)
    wf = f" smem wavefronts {v[4][0] / units:.0f} (ideal {v[4][1] / units:.0f})" if v[4][0] else ""
    print(f"line {ln:5d}: samples {100 * v[1] / max(tot_s, 1):4.1f} %  instr/unit {v[0] / units:7.1f} | {st}{wf}")

print("\n--- lines by shared-memory wavefronts ---")
tot_w = sum(v[4][0] for v in per.values())
print(f"total {tot_w / units:.0f} wavefronts per unit (ideal {sum(v[4][1] for v in per.values()) / units:.0f})")
for ln, v in sorted(per.items(), key=lambda kv: -kv[1][4][0])[:25]:
    if v[4][0]:
        ops = " ".join(f"{k} {c / units:.0f}" for k, c in sorted(v[2].items(), key=lambda kv: -kv[1])[:3])
        print(f"line {ln:5d}: {v[4][0] / units:7.0f} wavefronts (ideal {v[4][1] / units:6.0f})  | {ops}")
