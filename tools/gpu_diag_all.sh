#!/bin/bash
# Runs every diagnostic in its own process with a timeout; logs to gpurun_out/diag.log
mkdir -p gpurun_out
{
nvidia-smi --query-gpu=name,memory.total,clocks.sm --format=csv
for c in "gemm" "attn" "hifigan bf16" "hifigan fp16" "flow vosingle 1" "flow vosingle 0" "flow vomix 0"; do
  echo "=== $c"
  timeout 300 python tools/gpu_diag.py $c 2>&1 | grep -v "^sampling\|^two_cond" | tail -40
  echo "exit: ${PIPESTATUS[0]}"
done
} 2>&1 | tee gpurun_out/diag.log
