#!/bin/bash
mkdir -p gpurun_out
echo "=== all gpu tests ==="
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/tests.log 2>&1
tail -5 gpurun_out/tests.log
echo "=== train bench bf16 / fp32 ==="
timeout 600 python bench.py --workload train --steps 5 --warmup 3 > gpurun_out/train_bf16.json 2> gpurun_out/train_bf16.err
tail -3 gpurun_out/train_bf16.err; cat gpurun_out/train_bf16.json
timeout 600 python bench.py --workload train --train-precision fp32 --steps 3 --warmup 3 > gpurun_out/train_fp32.json 2> gpurun_out/train_fp32.err
cat gpurun_out/train_fp32.json
