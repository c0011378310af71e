set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q --timeout=300 -k "tcgen05 or tiny" > gpurun_out/t1.log 2>&1; echo "rc=$?" >> gpurun_out/t1.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; echo "rc=$?" >> gpurun_out/bench_ours.err
MSQ_GEMM_1CTA=1 timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_1cta.json 2> gpurun_out/bench_1cta.err; echo "rc=$?" >> gpurun_out/bench_1cta.err
CMD="python bench.py --steps 1 --warmup 1 --batch 32 --no-cpu-baseline"
timeout 600 $CMD > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu1.log 2>&1
