B4="python bench.py --config 4 --steps 1 --warmup 1 --no-cpu-baseline"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r2l_launches_config4.csv $B4 > gpurun_out/r2l_ncu1.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:dec_select_kernel -s 9 -c 9 -f -o gpurun_out/r2l_dec_select $B4 > gpurun_out/r2l_ncu2.log 2>&1
ls -la gpurun_out/
