set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout=300 -k "gemm_layernorm" > gpurun_out/t_ln.log 2>&1; echo "rc=$?" >> gpurun_out/t_ln.log
timeout 1500 python -m pytest tests -m gpu -x -q --timeout=900 > gpurun_out/t_all.log 2>&1; echo "rc=$?" >> gpurun_out/t_all.log
timeout 900 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; echo "rc=$?" >> gpurun_out/bench_ours.err
MSQ_NO_GEMM_LN=1 timeout 900 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/bench_noln.json 2> gpurun_out/bench_noln.err; echo "rc=$?" >> gpurun_out/bench_noln.err
