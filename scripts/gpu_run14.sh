set -x
mkdir -p gpurun_out
for b in 1 8 32 256; do
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --batch $b > gpurun_out/bench_b$b.json 2> gpurun_out/bench_b$b.err; echo "rc=$?" >> gpurun_out/bench_b$b.err
done
timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --batch 32 --precise > gpurun_out/bench_precise.json 2> gpurun_out/bench_precise.err; echo "rc=$?" >> gpurun_out/bench_precise.err
