#!/bin/bash
# One parametrised GPU launcher (replaces the per-experiment gpu_runNN.sh files): runs the given commands on the box,
# each under its own timeout, logging into gpurun_out/<tag>_<i>.log.   usage: scripts/gpu.sh <tag> '<cmd 1>' '<cmd 2>' ...
tag=$1; shift
mkdir -p gpurun_out
i=0
for cmd in "$@"; do
  i=$((i+1))
  echo "=== [$tag $i] $cmd"
  timeout ${GPU_CMD_TIMEOUT:-900} bash -c "$cmd" > gpurun_out/${tag}_${i}.log 2>&1
  echo "exit $?"; tail -n ${GPU_TAIL:-25} gpurun_out/${tag}_${i}.log
done
