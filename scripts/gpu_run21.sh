set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout=600 -rA -k "rn_tiny or resnet or full_size_vs_oracle or dropin_reproduces" 2>&1 | grep -v "^$" | tail -60 > gpurun_out/t_rn.log; echo "rc=$?" >> gpurun_out/t_rn.log
timeout 900 python bench.py --steps 6 --warmup 3 --backbone rn50 > gpurun_out/bench_rn50.json 2> gpurun_out/bench_rn50.err; echo "rc=$?" >> gpurun_out/bench_rn50.err
