#!/bin/bash
# float64 model kernels: parity tests + kernel-only timing (scripts/mlp_batch_bench.py), this library and, when
# present, the previous one (scripts/bin/libampc_prev.so)
cd "$(dirname "$0")/.."
out=gpurun_out
mkdir -p $out
timeout 300 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -6 > $out/s3m_gputest.log
timeout 100 python scripts/mlp_batch_bench.py > $out/s3m_mlp_default.txt 2>&1
[ -f scripts/bin/libampc_prev.so ] && timeout 100 python scripts/mlp_batch_bench.py scripts/bin/libampc_prev.so > $out/s3m_mlp_prev.txt 2>&1
timeout 100 python __graft_entry__.py --smoke 2>&1 | tail -2 > $out/s3m_smoke.log
tail -3 $out/s3m_gputest.log; cat $out/s3m_smoke.log; grep pred_diff $out/s3m_mlp_default.txt $out/s3m_mlp_prev.txt
