#!/bin/bash
# Multi-GPU lines (run through `gpurun --gpus N`): NCCL tests of the shards + batch / row / K shard bench lines.
# usage: bash tools/gpu_round2_multi.sh N
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2_multi_gpus_n$N.txt
( python -m pytest tests/test_gpu_shard.py tests/test_gpu_at_size.py -m gpu -q -k "nccl or two_devices" ) > gpurun_out/r2_pytest_multi_n$N.log 2>&1
cp gpurun_out/parity_report.json gpurun_out/r2_parity_report_multi_n$N.json 2>/dev/null
run() {  # name, extra args
  local name=$1; shift
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 8 --warmup 3 "$@" \
      > gpurun_out/r2_bench_${name}_n$N.json 2> gpurun_out/r2_bench_${name}_n$N.err
}
run batch
run row --shard row
run k_cfg4 --shard k --workload cfg4
run row_cfg5 --shard row --workload cfg5 --batch 8
run k --shard k
tail -3 gpurun_out/r2_pytest_multi_n$N.log
for f in gpurun_out/r2_bench_*_n$N.json; do echo $f; python -c "
import json,sys
try:
    d=json.load(open('$f')); print({k:d[k] for k in ('value','ms_per_step','n_gpus','scaling')}, d['config']['parallelism'])
except Exception as e: print('ERR', e)
"; done
