set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_dropin.py -m gpu -x -q --timeout=900 > gpurun_out/t_dropin.log 2>&1; echo "rc=$?" >> gpurun_out/t_dropin.log
