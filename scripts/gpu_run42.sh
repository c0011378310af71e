set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout=600 -k "invariance or config2 or tiny_golden or dropin_reproduces" > gpurun_out/t_chunk.log 2>&1; echo "rc=$?" >> gpurun_out/t_chunk.log
for i in 1 2; do
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_e2e$i.json 2> gpurun_out/bench_e2e$i.err; echo "rc=$?" >> gpurun_out/bench_e2e$i.err
done
