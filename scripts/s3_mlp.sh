#!/bin/bash
# float64 model kernels: kernel-only timing (scripts/mlp_batch_bench.py; also with the previous library when
# scripts/bin/libampc_prev.so is present), then the GPU parity suite and the smoke check
cd "$(dirname "$0")/.."
out=gpurun_out
mkdir -p $out
timeout 60 python scripts/mlp_batch_bench.py > $out/s3m_mlp_default.txt 2>&1
[ -f scripts/bin/libampc_prev.so ] && timeout 60 python scripts/mlp_batch_bench.py scripts/bin/libampc_prev.so > $out/s3m_mlp_prev.txt 2>&1
timeout 300 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -6 > $out/s3m_gputest.log
timeout 100 python __graft_entry__.py --smoke 2>&1 | tail -2 > $out/s3m_smoke.log
tail -3 $out/s3m_gputest.log; cat $out/s3m_smoke.log; cat $out/s3m_mlp_default.txt
