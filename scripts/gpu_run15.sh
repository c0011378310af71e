set -x
mkdir -p gpurun_out
timeout 1500 python scripts/agreement.py > gpurun_out/agreement.jsonl 2> gpurun_out/agreement.err; echo "rc=$?" >> gpurun_out/agreement.err
