set -x
mkdir -p gpurun_out
for n in 8 4 2; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${n}gpu.json 2> gpurun_out/bench_${n}gpu.err; echo "rc=$?" >> gpurun_out/bench_${n}gpu.err
done
