set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout=600 -rA -k "tiny_golden or full_size_vs_oracle" 2>&1 | grep -v "^$" | tail -40 > gpurun_out/t_fold.log; echo "rc=$?" >> gpurun_out/t_fold.log
for f in 1 0; do
MSQ_LNFOLD=$f timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_fold$f.json 2> gpurun_out/bench_fold$f.err; echo "rc=$?" >> gpurun_out/bench_fold$f.err
done
