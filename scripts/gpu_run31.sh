set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout=600 -k "invariance or tiny_golden or evaluate_batched or ragged" > gpurun_out/t_chunk.log 2>&1; echo "rc=$?" >> gpurun_out/t_chunk.log
timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_plan.json 2> gpurun_out/bench_plan.err; echo "rc=$?" >> gpurun_out/bench_plan.err
MSQ_CHUNK_MANUALS=32 timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c32.json 2> gpurun_out/bench_c32.err; echo "rc=$?" >> gpurun_out/bench_c32.err
timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --batch 50 > gpurun_out/bench_b50.json 2> gpurun_out/bench_b50.err
timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --batch 86 > gpurun_out/bench_b86.json 2> gpurun_out/bench_b86.err
