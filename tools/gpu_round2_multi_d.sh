#!/bin/bash
# usage: bash tools/gpu_round2_multi_d.sh N -- A/B of the fused FWD_B push (MPGCN_B200_SHARD_PUSH) on N GPUs
N=${1:-2}
mkdir -p gpurun_out
( python -m pytest tests/test_gpu_shard.py -m gpu -q -k "push or peer_exchange" ) > gpurun_out/r2_pytest_push.log 2>&1; tail -3 gpurun_out/r2_pytest_push.log
( MPGCN_B200_SHARD_PUSH=1 python -m pytest tests/test_gpu_shard.py -m gpu -q -k "nccl and row" ) > gpurun_out/r2_pytest_push_nccl.log 2>&1; tail -3 gpurun_out/r2_pytest_push_nccl.log
run() {  # name, extra args
  local name=$1; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 8 --warmup 3 --no-e2e "$@" \
      > gpurun_out/r2_ab_${name}_n$N.json 2> gpurun_out/r2_ab_${name}_n$N.err
}
MPGCN_B200_SHARD_PUSH=0 run pull --shard row
MPGCN_B200_SHARD_PUSH=1 run push --shard row
MPGCN_B200_SHARD_PUSH=1 run push_cfg5 --shard row --workload cfg5 --batch 8
MPGCN_B200_SHARD_PUSH=1 run push_cfg5_b1 --shard row --workload cfg5 --batch 1
for f in gpurun_out/r2_ab_*_n$N.json; do echo $f; python -c "
import json,sys
try:
    d=json.load(open('$f')); print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['roofline_step'].get('exchange_kernels_ms_per_step'), {k:round(v['ms']/d['steps'],2) for k,v in d['roofline']['per_stage'].items() if k in ('FWD_A','FWD_B','BWD_V','BWD_DX','LAYER_FWD','LAYER_BWD')})
except Exception as e: print('ERR', e)
"; tail -2 ${f%.json}.err; done
