set -x
mkdir -p gpurun_out
for cfg in "0 296" "0 444" "0 888" "0 1184" "3000 296" "6000 296"; do
set -- $cfg
MSQ_ATTN_STAGGER=$1 MSQ_ATTN_GRID=$2 timeout 300 python scripts/kernel_bench.py attn > gpurun_out/kb_attn_$1_$2.jsonl 2> gpurun_out/kb_attn.err
done
