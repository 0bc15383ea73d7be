#!/bin/bash
# float64 batch kernels: parity tests + kernel-only timing of both forms
cd "$(dirname "$0")/.."
out=gpurun_out
mkdir -p $out
timeout 300 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -6 > $out/s3m_gputest.log
AMPC_MLP_BLOCKED=1 timeout 100 python scripts/mlp_batch_bench.py > $out/s3m_mlp_blocked.txt 2>&1
AMPC_MLP_BLOCKED=0 timeout 100 python scripts/mlp_batch_bench.py > $out/s3m_mlp_one_sample.txt 2>&1
timeout 100 python scripts/mlp_batch_bench.py > $out/s3m_mlp_default.txt 2>&1
tail -3 $out/s3m_gputest.log; cat $out/s3m_mlp_blocked.txt $out/s3m_mlp_one_sample.txt $out/s3m_mlp_default.txt
