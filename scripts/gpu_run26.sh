set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout=600 -k "tiny_golden or full_size_vs_oracle" > gpurun_out/t_fold.log 2>&1; echo "rc=$?" >> gpurun_out/t_fold.log
timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_fold1.json 2> gpurun_out/bench_fold1.err; echo "rc=$?" >> gpurun_out/bench_fold1.err
timeout 600 python scripts/kernel_bench.py > gpurun_out/kernel_bench.jsonl 2> gpurun_out/kernel_bench.err
