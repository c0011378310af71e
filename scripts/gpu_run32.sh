set -x
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --batch 32 --no-cpu-baseline"
timeout 600 $CMD > gpurun_out/plain.log 2>&1 || exit 1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r1h_launches_bench_b32.csv $CMD > gpurun_out/ncu_l.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 60 -c 8 -f -o gpurun_out/prof_gemm_tc_h $CMD > gpurun_out/ncu_g.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attention_tc_kernel -s 30 -c 2 -f -o gpurun_out/prof_attention_h $CMD > gpurun_out/ncu_a.log 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q --timeout=900 -k "long_manual or config2" > gpurun_out/t_new.log 2>&1; echo "rc=$?" >> gpurun_out/t_new.log
