#!/bin/bash
# First GPU bring-up: safe kernels first, the tcgen05 kernel in its own process under a timeout.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/smi.txt 2>&1
echo "=== postproc ===" 
timeout 600 python -m pytest tests/test_gpu_postproc.py -q -m gpu -x 2>&1 | tail -25 | tee gpurun_out/t_postproc.log
echo "=== fp32 conv ==="
timeout 600 python -m pytest tests/test_gpu_conv.py -q -m gpu -k "fp32 or conv1" 2>&1 | tail -25 | tee gpurun_out/t_conv_fp32.log
echo "=== tc conv (one case) ==="
timeout 120 python -m pytest tests/test_gpu_conv.py -q -m gpu -s -k "bf16 and B2_18x18_1024-512_k1s1" 2>&1 | tail -30 | tee gpurun_out/t_conv_tc1.log
echo "=== tc conv (all) ==="
timeout 600 python -m pytest tests/test_gpu_conv.py -q -m gpu -s -k "bf16 or linearity" 2>&1 | tail -60 | tee gpurun_out/t_conv_tc.log
echo "=== network ==="
timeout 900 python -m pytest tests/test_gpu_network.py -q -m gpu -s 2>&1 | tail -40 | tee gpurun_out/t_net.log
