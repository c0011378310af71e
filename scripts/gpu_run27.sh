set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q --timeout=600 -k "deferred" > gpurun_out/t_def.log 2>&1; echo "rc=$?" >> gpurun_out/t_def.log
timeout 600 python scripts/kernel_bench.py deferred > gpurun_out/kb_deferred.jsonl 2> gpurun_out/kb_deferred.err
