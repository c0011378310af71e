set -x
mkdir -p gpurun_out
for c in 2 4 8 16 64; do
MSQ_CHUNK_MANUALS=$c timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_chunk$c.json 2> gpurun_out/bench_chunk$c.err; echo "rc=$?" >> gpurun_out/bench_chunk$c.err
done
