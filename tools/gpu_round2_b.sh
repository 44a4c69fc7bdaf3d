#!/bin/bash
# Round-2 profile collection on ONE B200 (run through gpurun from the repo root):
#   ncu launch list of the bench command, --set full captures of the pair kernel / 1-CTA kernel / LSTM kernels
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2_launches.csv \
    python bench.py --steps 1 --warmup 1 --profile --batch 4 > gpurun_out/r2_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:contract2_kernel --launch-skip 24 -c 4 -o gpurun_out/r2_prof_contract2 -f \
    python bench.py --steps 1 --warmup 1 --profile --batch 4 > gpurun_out/r2_prof_contract2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:contract_kernel --launch-skip 18 -c 4 -o gpurun_out/r2_prof_mix -f \
    python bench.py --steps 1 --warmup 1 --profile --batch 4 > gpurun_out/r2_prof_mix.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lstm_ -c 2 -o gpurun_out/r2_prof_lstm -f \
    python profiles/lstm_probe.py > gpurun_out/r2_prof_lstm.log 2>&1
ls -la gpurun_out/*.ncu-rep
python tools/small_n_bench.py > gpurun_out/r2_small_n.jsonl 2> gpurun_out/r2_small_n.err
