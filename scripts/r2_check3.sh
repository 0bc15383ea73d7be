#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ilqr_gpu.py tests/test_thresh_gpu.py tests/test_closed_loop_gpu.py -m gpu -q --timeout 600 --tb=short 2>&1 | tail -30 > gpurun_out/pytest_ilqr.log
timeout 120 python scripts/ilqr_one.py > gpurun_out/ilqr_prof.json 2> gpurun_out/ilqr_prof.err
timeout 300 python bench.py --workload c4 --steps 20 --warmup 3 > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err
timeout 300 python bench.py --workload c5 --steps 5 --warmup 1 > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err
tail -8 gpurun_out/pytest_ilqr.log; cat gpurun_out/ilqr_prof.json; tail -3 gpurun_out/ilqr_prof.err
for f in c4 c5; do echo "== $f"; cut -c1-260 gpurun_out/bench_$f.json; tail -3 gpurun_out/bench_$f.err; done
