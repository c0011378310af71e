set -x
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --batch 32 --no-cpu-baseline"
timeout 600 $CMD > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_ln -s 30 -c 2 -o gpurun_out/prof_gemm_ln $CMD > gpurun_out/ncu2.log 2>&1
