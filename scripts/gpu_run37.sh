set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q --timeout=900 -rA -k "wired_production" 2>&1 | tail -25 > gpurun_out/t_prod.log; echo "rc=$?" >> gpurun_out/t_prod.log
timeout 900 python bench.py --steps 4 --warmup 3 --text roberta-large --backbone rn50 --batch 32 > gpurun_out/bench_prod.json 2> gpurun_out/bench_prod.err; echo "rc=$?" >> gpurun_out/bench_prod.err
