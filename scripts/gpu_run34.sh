set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout=600 -k "rn_tiny or resnet or rn50" > gpurun_out/t_rn.log 2>&1; echo "rc=$?" >> gpurun_out/t_rn.log
timeout 600 python bench.py --steps 6 --warmup 3 --backbone rn50 --no-cpu-baseline > gpurun_out/bench_rn50.json 2> gpurun_out/bench_rn50.err; echo "rc=$?" >> gpurun_out/bench_rn50.err
