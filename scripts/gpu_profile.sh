#!/bin/bash
# ncu evidence for the bench command (1 GPU).  Run under gpurun from the repo root.
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:mppi_rollout -s 4 -c 2 -f -o gpurun_out/prof \
    python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_under_ncu_full.log 2>&1
ls -la gpurun_out
