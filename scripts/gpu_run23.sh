set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q --timeout=600 -k "mm_tiny or full_size_vs_oracle" > gpurun_out/t_slab0.log 2>&1; echo "rc=$?" >> gpurun_out/t_slab0.log
for s in 0 12544 6272 25088; do
MSQ_SLAB_ROWS=$s timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_slab$s.json 2> gpurun_out/bench_slab$s.err; echo "rc=$?" >> gpurun_out/bench_slab$s.err
done
MSQ_SLAB_ROWS=12544 timeout 600 python -m pytest tests -m gpu -x -q --timeout=600 -k "mm_tiny or full_size_vs_oracle or invariance" > gpurun_out/t_slab1.log 2>&1; echo "rc=$?" >> gpurun_out/t_slab1.log
