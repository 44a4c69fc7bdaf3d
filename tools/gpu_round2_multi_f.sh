#!/bin/bash
# 4 GPUs: the multi-GPU tests once more on the final code + hybrid (row 2 x batch 2) lines
mkdir -p gpurun_out
( python -m pytest tests/test_gpu_shard.py tests/test_gpu_at_size.py -m gpu -q -k "nccl or two_devices" ) > gpurun_out/r2_pytest_multif.log 2>&1; tail -3 gpurun_out/r2_pytest_multif.log
cp gpurun_out/parity_report.json gpurun_out/r2_parity_report_multi_final.json 2>/dev/null
bash tools/gpu_round2_multi_e.sh 4 2>&1 | grep -v "^\*\*\*\|NCCL version" | head -20
