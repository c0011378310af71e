set -x
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --batch 32 --no-cpu-baseline"
timeout 600 $CMD > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 62 -c 8 -o gpurun_out/prof_gemm_tc $CMD > gpurun_out/ncu2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attention_tc -s 13 -c 2 -o gpurun_out/prof_attention $CMD > gpurun_out/ncu3.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:layernorm -s 30 -c 1 -o gpurun_out/prof_layernorm $CMD > gpurun_out/ncu4.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:beam_search -c 1 -o gpurun_out/prof_beam $CMD > gpurun_out/ncu5.log 2>&1
ls -la gpurun_out
