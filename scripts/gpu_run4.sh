set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q --timeout=300 -k "attention or tcgen05" > gpurun_out/t1.log 2>&1; echo "rc=$?" >> gpurun_out/t1.log
timeout 900 python -m pytest tests -m gpu -q --timeout=400 -k "tiny" -s > gpurun_out/t2.log 2>&1; echo "rc=$?" >> gpurun_out/t2.log
timeout 1200 python -m pytest tests -m gpu -q --timeout=900 -k "full_size" -s > gpurun_out/t3.log 2>&1; echo "rc=$?" >> gpurun_out/t3.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; echo "rc=$?" >> gpurun_out/bench_ours.err
MSQ_GEMM_1CTA=1 timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_1cta.json 2> gpurun_out/bench_1cta.err; echo "rc=$?" >> gpurun_out/bench_1cta.err
