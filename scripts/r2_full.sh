#!/bin/bash
# round 2: full GPU validation + the judged artefacts (1 GPU)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
timeout 1800 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_default.json 2> gpurun_out/bench_ref_default.err
timeout 300 python bench.py --precision bf16 --no-cpu --steps 200 --warmup 20 > gpurun_out/bench_bf16.json 2> gpurun_out/bench_bf16.err
timeout 300 python bench.py --precision fp32 --no-cpu --steps 10 --warmup 3 > gpurun_out/bench_fp32.json 2> gpurun_out/bench_fp32.err
timeout 300 python bench.py --workload c2 --steps 200 --warmup 20 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:mppi_rollout -s 4 -c 2 -f -o gpurun_out/prof_r02 \
    python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_under_ncu_full.log 2>&1
tail -6 gpurun_out/pytest_gpu.log; tail -5 gpurun_out/smoke.log
for f in default ref_default bf16 fp32 c2; do echo "== $f"; cut -c1-330 gpurun_out/bench_$f.json; tail -2 gpurun_out/bench_$f.err; done
