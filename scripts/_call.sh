mkdir -p gpurun_out
N=${NGPU:-4}
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --config l-voc-train --steps 6 --warmup 3 > gpurun_out/r02_bench_l_voc_train_n$N.json 2> gpurun_out/r02_bench_l_voc_train_n$N.err
tail -c 200 gpurun_out/r02_bench_l_voc_train_n$N.err
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --no-cpu-baseline --no-gpu-reference > gpurun_out/r02_bench_m_n$N.json 2> gpurun_out/r02_bench_m_n$N.err
tail -c 200 gpurun_out/r02_bench_m_n$N.err
for f in gpurun_out/r02_bench_*_n$N.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
for k in ("metric","n_gpus","value","ms_per_step","e2e","e2e_uint8_frames","whole_box"):
    if k in d: print(k, d[k])
PY
done
