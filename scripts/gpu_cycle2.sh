#!/bin/bash
mkdir -p gpurun_out
echo "=== conv tests ==="
timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_network.py -q -m gpu -x 2>&1 | tail -5
echo "=== A/B ==="
timeout 900 python scripts/ab_layers.py 2>&1 | tail -90 | tee gpurun_out/ab_layers.txt
