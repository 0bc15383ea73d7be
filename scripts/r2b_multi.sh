#!/bin/bash
# round 2 (dz kernel) multi-GPU run: bash scripts/r2b_multi.sh N   (under gpurun --gpus N)
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29511 scripts/multi_gpu_check.py > gpurun_out/r02b_multi_gpu_check_n$N.log 2>&1
timeout 300 $TR --master-port 29512 bench.py --gpus $N --steps 200 --warmup 20 > gpurun_out/r02b_bench_strong_n$N.json 2> gpurun_out/bench_strong_n$N.err
timeout 300 $TR --master-port 29513 bench.py --gpus $N --steps 200 --warmup 20 --scaling weak > gpurun_out/r02b_bench_weak_n$N.json 2> gpurun_out/bench_weak_n$N.err
timeout 300 $TR --master-port 29514 bench.py --gpus $N --steps 5 --warmup 1 --workload c5 > gpurun_out/r02b_bench_c5_n$N.json 2> gpurun_out/bench_c5_n$N.err
tail -3 gpurun_out/r02b_multi_gpu_check_n$N.log
for f in strong weak c5; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02b_bench_${f}_n$N.json").read().strip().splitlines()[-1])
    print("$f", {k:d.get(k) for k in ("value","ms_per_step","n_gpus","scaling")}, "e2e", d["e2e"]["value"], "parity", d.get("parity",{}).get("max_abs_du_vs_single_gpu"))
except Exception as e:
    print("ERR $f", e)
PY
done
