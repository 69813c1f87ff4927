mkdir -p gpurun_out
timeout 420 python bench.py > gpurun_out/r02_bench_m.json 2> gpurun_out/r02_bench_m.err
tail -c 300 gpurun_out/r02_bench_m.err
timeout 300 python bench.py --config s-city --no-cpu-baseline > gpurun_out/r02_bench_s_city.json 2> gpurun_out/r02_bench_s_city.err
timeout 400 python bench.py --config s-camvid --sweep --no-cpu-baseline > gpurun_out/r02_bench_s_camvid.json 2> gpurun_out/r02_bench_s_camvid.err
python - <<'PY'
import json
for c in ("m","s_city","s_camvid"):
    try:
        d=json.loads(open(f"gpurun_out/r02_bench_{c}.json").read().strip().splitlines()[-1])
    except Exception as e:
        print(c,"ERR",e); print(open(f"gpurun_out/r02_bench_{c}.err").read()[-800:]); continue
    print("==",c,"value",round(d["value"],1),"ms",round(d["ms_per_step"],3),"e2e",round(d["e2e"]["value"],1),"u8",round(d.get("e2e_uint8_frames",{}).get("value",0),1),"roofline",round(d["roofline"]["frac"],4),"clocks",d["clocks"].get("sm_mhz"),d["clocks"].get("samples"),"launches",d.get("gpu_launches_per_step"))
    print("   patch_conv",d.get("patch_conv"),"heads",d.get("heads"))
    for k,v in d["kernels"].items(): print("   ",k,round(v["ms"]*1e3,1),"us",round(v["frac"],3),v.get("kernel",""))
    if d.get("cpu_baseline"): print("   cpu",d["cpu_baseline"])
PY
