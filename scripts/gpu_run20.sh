set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout=600 -k "rn or resnet" > gpurun_out/t_rn.log 2>&1; echo "rc=$?" >> gpurun_out/t_rn.log
timeout 1200 python -m pytest tests -m gpu -x -q --timeout=900 -k "not rn and not resnet" > gpurun_out/t_rest.log 2>&1; echo "rc=$?" >> gpurun_out/t_rest.log
