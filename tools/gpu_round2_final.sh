#!/bin/bash
# Final single-GPU pass of the round: the full GPU suite, smoke(), the default bench line, the reference arm, cfg5 baselines for the shard curves.
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q --durations=15 ) > gpurun_out/r2_pytest_final.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_pytest_final.log
cp gpurun_out/parity_report.json gpurun_out/r2_parity_report_final.json 2>/dev/null
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2_smoke.log
( time python bench.py --steps 20 --warmup 5 ) > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err
( time python bench.py --impl reference --steps 5 --warmup 1 ) > gpurun_out/r2_bench_ref_final.json 2> gpurun_out/r2_bench_ref_final.err
MPGCN_B200_BRANCH_STREAMS=1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline > gpurun_out/r2_bench_branchstreams.json 2> gpurun_out/r2_bench_branchstreams.err
( MPGCN_B200_BRANCH_STREAMS=1 python -m pytest tests/test_gpu_parity.py tests/test_gpu_at_size.py -m gpu -q -k "full_model or training_is_equivalent or trainer_call" ) > gpurun_out/r2_pytest_branchstreams.log 2>&1; tail -2 gpurun_out/r2_pytest_branchstreams.log
python bench.py --workload cfg5 --batch 8 --steps 4 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-e2e > gpurun_out/r2_bench_cfg5_b8.json 2> gpurun_out/r2_bench_cfg5_b8.err
python bench.py --workload cfg5 --batch 1 --steps 6 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-e2e > gpurun_out/r2_bench_cfg5_b1.json 2> gpurun_out/r2_bench_cfg5_b1.err
tail -4 gpurun_out/r2_pytest_final.log; cat gpurun_out/r2_smoke.log | tail -3
for f in r2_bench_final r2_bench_branchstreams r2_bench_ref_final r2_bench_cfg5_b8 r2_bench_cfg5_b1; do python -c "
import json
try:
    d=json.load(open('gpurun_out/$f.json')); print('$f', {k:d[k] for k in ('value','ms_per_step')})
except Exception as e: print('$f ERR', e)
"; done
