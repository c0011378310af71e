set -x
mkdir -p gpurun_out
timeout 300 python scripts/kernel_bench.py deferred > gpurun_out/plain.log 2>&1 || exit 1
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'gemm_tc_kernel<float.*2>' -s 4 -c 1 -o gpurun_out/prof_resln -f python scripts/kernel_bench.py deferred > gpurun_out/ncu_resln.log 2>&1
ls -la gpurun_out/*.ncu-rep
