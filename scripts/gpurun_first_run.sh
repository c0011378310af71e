set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
ls /root/reference > gpurun_out/ref_ls.txt 2>&1
timeout 600 python -m pytest tests -m gpu -x -q --timeout=300 -k "ffma or layernorm or attention" > gpurun_out/t1_kernels.log 2>&1; echo "rc=$?" >> gpurun_out/t1_kernels.log
timeout 600 python -m pytest tests -m gpu -x -q --timeout=300 -k "decode" > gpurun_out/t2_decode.log 2>&1; echo "rc=$?" >> gpurun_out/t2_decode.log
timeout 600 python -m pytest tests -m gpu -q --timeout=300 -k "tcgen05" > gpurun_out/t3_tc.log 2>&1; echo "rc=$?" >> gpurun_out/t3_tc.log
timeout 900 python -m pytest tests -m gpu -q --timeout=400 -k "tiny and True" -s > gpurun_out/t4_tiny_fp32.log 2>&1; echo "rc=$?" >> gpurun_out/t4_tiny_fp32.log
timeout 900 python -m pytest tests -m gpu -q --timeout=400 -k "tiny and False" -s > gpurun_out/t5_tiny_bf16.log 2>&1; echo "rc=$?" >> gpurun_out/t5_tiny_bf16.log
timeout 600 python __graft_entry__.py smoke > gpurun_out/t6_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/t6_smoke.log
tail -5 gpurun_out/t*.log
