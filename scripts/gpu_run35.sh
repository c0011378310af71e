set -x
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 1 --batch 1 --no-cpu-baseline"
timeout 600 $CMD > gpurun_out/bench_b1.json 2> gpurun_out/bench_b1.err || exit 1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_b1.csv $CMD > gpurun_out/ncu_b1.log 2>&1
