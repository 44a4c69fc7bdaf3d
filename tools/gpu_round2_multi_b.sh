#!/bin/bash
# usage: bash tools/gpu_round2_multi_b.sh N   -- NCCL tests + row-shard lines (peer-memory exchange vs NCCL exchange)
N=${1:-2}
mkdir -p gpurun_out
( python -m pytest tests/test_gpu_shard.py tests/test_gpu_at_size.py -m gpu -q -k "nccl or two_devices" ) > gpurun_out/r2_pytest_multib_n$N.log 2>&1
cp gpurun_out/parity_report.json gpurun_out/r2_parity_report_multib_n$N.json 2>/dev/null
run() {  # name, env, extra args
  local name=$1; shift
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 8 --warmup 3 --no-e2e "$@" \
      > gpurun_out/r2_bench_${name}_n$N.json 2> gpurun_out/r2_bench_${name}_n$N.err
}
run rowpeer --shard row
MPGCN_B200_SHARD_EXCHANGE=nccl run rownccl --shard row
run rowpeer_cfg5 --shard row --workload cfg5 --batch 8
tail -5 gpurun_out/r2_pytest_multib_n$N.log
for f in gpurun_out/r2_bench_row*_n$N.json; do echo $f; python -c "
import json,sys
try:
    d=json.load(open('$f')); print({k:d[k] for k in ('value','ms_per_step','n_gpus','scaling')}, d['config']['parallelism'], d.get('shard',{}).get('exchange'))
except Exception as e: print('ERR', e)
"; tail -2 ${f%.json}.err; done
