#!/bin/bash
# usage: bash tools/gpu_round2_multi_c.sh N   -- scaling lines on N GPUs (+ the NCCL / peer shard tests when N == 2)
N=${1:-2}
mkdir -p gpurun_out
if [ "$N" = "2" ]; then
  ( python -m pytest tests/test_gpu_shard.py tests/test_gpu_at_size.py -m gpu -q -k "nccl or two_devices" ) > gpurun_out/r2_pytest_multic_n$N.log 2>&1
  cp gpurun_out/parity_report.json gpurun_out/r2_parity_report_multi.json 2>/dev/null
  tail -3 gpurun_out/r2_pytest_multic_n$N.log
fi
run() {  # name, extra args
  local name=$1; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 8 --warmup 3 "$@" \
      > gpurun_out/r2_final_${name}_n$N.json 2> gpurun_out/r2_final_${name}_n$N.err
}
run batch
run row --shard row
run row_cfg5 --shard row --workload cfg5 --batch 8
run k_cfg4 --shard k --workload cfg4
run row_cfg5_b1 --shard row --workload cfg5 --batch 1
for f in gpurun_out/r2_final_*_n$N.json; do echo $f; python -c "
import json,sys
try:
    d=json.load(open('$f')); print({k:d[k] for k in ('value','ms_per_step','n_gpus','scaling')}, d['config']['parallelism'], (d.get('shard') or {}).get('exchange'), d['roofline_step'].get('exchange_kernels_ms_per_step'))
except Exception as e: print('ERR', e)
"; tail -2 ${f%.json}.err; done
