set -x
mkdir -p gpurun_out
timeout 600 python bench.py --steps 1 --warmup 1 --batch 32 --no-cpu-baseline > gpurun_out/b32.json 2> gpurun_out/b32.err || exit 1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_fold1.csv python bench.py --steps 1 --warmup 1 --batch 32 --no-cpu-baseline > gpurun_out/ncu1.log 2>&1
