#!/bin/bash
# Collects the round's profile artefacts on a B200 box (run through gpurun from the repo root):
#   bash profiles/collect.sh        -> gpurun_out/{bench_n1.json, launches.csv, prof_*.ncu-rep, parity_report.json}
# then, here:  python profiles/summarize.py launches gpurun_out/launches.csv > profiles/launches_r1.txt   etc.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --timeout 600 2>&1 | tail -2
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:contract2_kernel --launch-skip 48 -c 4 -o gpurun_out/prof_contract2 -f \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/prof_contract2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:contract_kernel --launch-skip 36 -c 4 -o gpurun_out/prof_mix -f \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/prof_mix.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lstm_ -c 2 -o gpurun_out/prof_lstm -f \
    python profiles/lstm_probe.py > gpurun_out/prof_lstm.log 2>&1
ls -la gpurun_out/*.ncu-rep
