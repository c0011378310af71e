set -x
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --batch 32 --no-cpu-baseline --backbone rn50"
timeout 600 $CMD > gpurun_out/plain_rn.log 2>&1 || exit 1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r1h_launches_rn50_b32.csv $CMD > gpurun_out/ncu_lrn.log 2>&1
