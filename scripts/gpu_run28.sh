set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q --timeout=600 -k "deferred or gemm_tcgen05 or full_size_vs_oracle" > gpurun_out/t_def.log 2>&1; echo "rc=$?" >> gpurun_out/t_def.log
timeout 600 python scripts/kernel_bench.py deferred > gpurun_out/kb_deferred.jsonl 2> gpurun_out/kb_deferred.err
timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_fold1.json 2> gpurun_out/bench_fold1.err; echo "rc=$?" >> gpurun_out/bench_fold1.err
