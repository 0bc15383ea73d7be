#!/bin/bash
# round 2: iLQR v2 parity + timing, new bench workloads
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ilqr_gpu.py tests/test_thresh_gpu.py -m gpu -q --timeout 600 --tb=short 2>&1 | tail -30 > gpurun_out/pytest_ilqr.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1
timeout 300 python bench.py --workload c4 --steps 20 --warmup 3 > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err
AMPC_ILQR_NO_SMEM=1 timeout 300 python bench.py --workload c4 --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_c4_nosmem.json 2> gpurun_out/bench_c4_nosmem.err
timeout 300 python bench.py --workload c5 --steps 5 --warmup 1 > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err
timeout 300 python bench.py --workload c2 --steps 100 --warmup 10 --no-cpu > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err
timeout 600 python bench.py --steps 100 --warmup 10 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err
tail -12 gpurun_out/pytest_ilqr.log; tail -4 gpurun_out/smoke.log
for f in c4 c4_nosmem c5 c2 c3; do echo "== $f"; cut -c1-260 gpurun_out/bench_$f.json; tail -3 gpurun_out/bench_$f.err; done
