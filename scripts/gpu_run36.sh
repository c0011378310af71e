set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --timeout=900 > gpurun_out/t_all.log 2>&1; echo "rc=$?" >> gpurun_out/t_all.log
timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; echo "rc=$?" >> gpurun_out/bench_ours.err
timeout 600 python bench.py --steps 8 --warmup 3 --batch 1 --no-cpu-baseline > gpurun_out/bench_b1.json 2> gpurun_out/bench_b1.err
