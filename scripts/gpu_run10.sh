set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --timeout=900 -k "edge or rejects or empty or roberta" > gpurun_out/t_edge.log 2>&1; echo "rc=$?" >> gpurun_out/t_edge.log
