#!/bin/bash
# Final-state evidence of a round (run on the GPU box through gpurun): the driver-like default bench line, the ncu launch list
# of the same command at batch 32, `ncu --set full` captures of the top kernels of the headline and of the fine-tuning step.
# usage: scripts/capture_final.sh <tag>
tag=${1:-final}
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 1 --batch 32 --no-cpu-baseline --no-extra-configs --no-parity"
set -x
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err || exit 1
$B > gpurun_out/${tag}_b32.json 2>&1 || exit 1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2400 --csv --log-file gpurun_out/${tag}_launches.csv $B > gpurun_out/${tag}_ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 100 -c 8 -f -o gpurun_out/${tag}_gemm_x3 $B > gpurun_out/${tag}_ncu2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_ws_kernel -s 30 -c 2 -f -o gpurun_out/${tag}_attn_ws $B > gpurun_out/${tag}_ncu3.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:attention_bwd_dq_mma_kernel|ln_bwd_kernel|act_bwd_kernel|attention_bwd_dkv_tc_kernel" -s 8 -c 8 -f -o gpurun_out/${tag}_finetune python scripts/profile_finetune.py 8 > gpurun_out/${tag}_ncu4.log 2>&1
ls -la gpurun_out/${tag}_*
