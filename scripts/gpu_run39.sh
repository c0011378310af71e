set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout=600 -k "attention or tiny_golden or full_size_vs_oracle" > gpurun_out/t_attn.log 2>&1; echo "rc=$?" >> gpurun_out/t_attn.log
timeout 600 python scripts/kernel_bench.py attn > gpurun_out/kb_attn.jsonl 2> gpurun_out/kb_attn.err
timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; echo "rc=$?" >> gpurun_out/bench_ours.err
