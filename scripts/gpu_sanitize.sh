#!/bin/bash
# compute-sanitizer memcheck + synccheck over a small forward (default and forced CTA-pair plans), a training step,
# the I/O kernels and the multi-batch NMS.  Logs -> gpurun_out/r2_sanitizer_*.log
mkdir -p gpurun_out
for tool in memcheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitize_small.py > gpurun_out/r2_sanitizer_$tool.log 2>&1
  echo "== $tool: exit $?"; grep -E "ok|ERROR SUMMARY|Error|error" gpurun_out/r2_sanitizer_$tool.log | head -12
done
