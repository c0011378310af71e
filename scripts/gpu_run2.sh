set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout=900 -k "full_size" -s > gpurun_out/t7_full.log 2>&1; echo "rc=$?" >> gpurun_out/t7_full.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; echo "rc=$?" >> gpurun_out/bench_ours.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "rc=$?" >> gpurun_out/bench_ref.err
CMD="python bench.py --steps 1 --warmup 1 --batch 32 --no-cpu-baseline"
timeout 600 $CMD > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 40 -c 3 -o gpurun_out/prof_gemm_tc $CMD > gpurun_out/ncu2.log 2>&1
ls -la gpurun_out
