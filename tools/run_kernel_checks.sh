#!/usr/bin/env bash
# Run every kernel_check case in its own process with a timeout; logs under gpurun_out/kcheck/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/kcheck
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/kcheck/gpu.txt 2>&1
cases="${*:-$(python tools/kernel_check.py list)}"
for c in $cases; do
  timeout 300 python tools/kernel_check.py "$c" > "gpurun_out/kcheck/$c.log" 2>&1
  echo "case $c exit $?" | tee -a gpurun_out/kcheck/summary.txt
  grep -E "^\[(OK|FAIL)\]|^==|Error|error|timeout|gemm M=|attn causal" "gpurun_out/kcheck/$c.log" | tail -40
done
