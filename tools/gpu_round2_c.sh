#!/bin/bash
mkdir -p gpurun_out
( time python -m pytest tests/test_gpu_at_size.py -m gpu -q -k "layer_matches_oracle_at_size" ) > gpurun_out/r2_pytest2.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_pytest2.log
cp gpurun_out/parity_report.json gpurun_out/r2_parity_report2.json 2>/dev/null
python bench.py --workload cfg5 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_cfg5.json 2> gpurun_out/r2_bench_cfg5.err
python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_headline2.json 2> gpurun_out/r2_bench_headline2.err
bash tools/gpu_round2_b.sh > gpurun_out/r2_profile.log 2>&1
tail -3 gpurun_out/r2_pytest2.log
