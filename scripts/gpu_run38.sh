set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_dropin.py -m gpu -x -q --timeout=600 > gpurun_out/t_dropin.log 2>&1; echo "rc=$?" >> gpurun_out/t_dropin.log
