#!/bin/bash
# usage: bash tools/gpu_round2_multi_e.sh N -- hybrid row x batch shard lines on N GPUs
N=${1:-8}
mkdir -p gpurun_out
run() {  # name, extra args
  local name=$1; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 8 --warmup 3 "$@" \
      > gpurun_out/r2_hyb_${name}_n$N.json 2> gpurun_out/r2_hyb_${name}_n$N.err
}
run row2 --shard row --row-ranks 2
[ "$N" -ge 8 ] && run row4 --shard row --row-ranks 4
run row2_cfg5 --shard row --row-ranks 2 --workload cfg5 --batch 8
[ "$N" -ge 8 ] && run row4_cfg5 --shard row --row-ranks 4 --workload cfg5 --batch 8
for f in gpurun_out/r2_hyb_*_n$N.json; do echo $f; python -c "
import json,sys
try:
    d=json.load(open('$f')); print({k:d[k] for k in ('value','ms_per_step','n_gpus','scaling')}, d['config']['parallelism'], (d.get('shard') or {}).get('exchange'), d['roofline_step'].get('exchange_kernels_ms_per_step'))
except Exception as e: print('ERR', e)
"; tail -2 ${f%.json}.err; done
