#!/bin/bash
# round 2, first GPU round trip: parity tests, precision measurement, bench lines for both tensor-core modes
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -x 2>&1 | tail -40 > gpurun_out/pytest_gpu.log
timeout 900 python scripts/measure_precision.py > gpurun_out/precision.jsonl 2> gpurun_out/precision.err
timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu --precision bf16 > gpurun_out/bench_bf16.json 2> gpurun_out/bench_bf16.err
timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu --precision fp16 > gpurun_out/bench_fp16.json 2> gpurun_out/bench_fp16.err
tail -15 gpurun_out/pytest_gpu.log; cat gpurun_out/precision.jsonl | tail -40; tail -3 gpurun_out/precision.err; cat gpurun_out/bench_bf16.json gpurun_out/bench_fp16.json | cut -c1-400
