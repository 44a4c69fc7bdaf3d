#!/bin/bash
mkdir -p gpurun_out
python tools/lstm_poly_probe.py > gpurun_out/r2_lstm_poly.jsonl 2> gpurun_out/r2_lstm_poly.err
cat gpurun_out/r2_lstm_poly.jsonl
MPGCN_B200_LSTM_POLY=1 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "lstm" > gpurun_out/r2_pytest_lstm_poly1.log 2>&1; tail -2 gpurun_out/r2_pytest_lstm_poly1.log
MPGCN_B200_LSTM_POLY=2 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "lstm" > gpurun_out/r2_pytest_lstm_poly2.log 2>&1; tail -2 gpurun_out/r2_pytest_lstm_poly2.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:contract2_kernel --launch-skip 24 -c 4 -o gpurun_out/r2_prof_contract2 -f \
    python bench.py --steps 1 --warmup 1 --profile --batch 4 > gpurun_out/r2_prof_contract2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:contract_kernel --launch-skip 18 -c 4 -o gpurun_out/r2_prof_mix -f \
    python bench.py --steps 1 --warmup 1 --profile --batch 4 > gpurun_out/r2_prof_mix.log 2>&1
ls -la gpurun_out/r2_*.ncu-rep
