set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --timeout=900 > gpurun_out/t_all.log 2>&1; echo "rc=$?" >> gpurun_out/t_all.log
timeout 600 python bench.py --steps 8 --warmup 3 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; echo "rc=$?" >> gpurun_out/bench_ours.err
timeout 600 python bench.py --steps 6 --warmup 3 --backbone rn50 --no-cpu-baseline > gpurun_out/bench_rn50.json 2> gpurun_out/bench_rn50.err; echo "rc=$?" >> gpurun_out/bench_rn50.err
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.log 2>&1; echo "rc=$?" >> gpurun_out/smoke.log
