mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2_gpus.txt 2>&1
nproc >> gpurun_out/r2_gpus.txt
( time python -m pytest tests -m gpu -q --durations=25 ) > gpurun_out/r2_pytest1.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_pytest1.log
cp gpurun_out/parity_report.json gpurun_out/r2_parity_report1.json 2>/dev/null
( time python bench.py --steps 10 --warmup 3 ) > gpurun_out/r2_bench_headline1.json 2> gpurun_out/r2_bench_headline1.err
( time python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/r2_bench_ref1.json 2> gpurun_out/r2_bench_ref1.err
for w in cfg2 cfg3 cfg4 cfg5; do
  python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_$w.json 2> gpurun_out/r2_bench_$w.err
done
tail -3 gpurun_out/r2_pytest1.log
