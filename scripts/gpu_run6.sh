set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --timeout=900 > gpurun_out/t_all.log 2>&1; echo "rc=$?" >> gpurun_out/t_all.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "rc=$?" >> gpurun_out/smoke.log
