set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout=300 -k "mm_tiny_golden or mm_rn_tiny_golden" > gpurun_out/plain_tiny.log 2>&1 || exit 1
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout=1400 -k "mm_tiny_golden or mm_rn_tiny_golden" > gpurun_out/sanitizer_memcheck.log 2>&1; echo "rc=$?" >> gpurun_out/sanitizer_memcheck.log
