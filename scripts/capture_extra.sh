set -x
B4="python bench.py --config 4 --steps 1 --warmup 1 --no-cpu-baseline"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r2k_launches_config4.csv $B4 > gpurun_out/r2k_ncu5.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dec_select_kernel -s 12 -c 3 -f -o gpurun_out/r2k_dec_select $B4 > gpurun_out/r2k_ncu6.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 60 -c 8 -f -o gpurun_out/r2k_finetune_fwd_gemm python scripts/profile_finetune.py 8 > gpurun_out/r2k_ncu7.log 2>&1
ls -la gpurun_out/r2k_dec_select.ncu-rep gpurun_out/r2k_finetune_fwd_gemm.ncu-rep gpurun_out/r2k_launches_config4.csv
