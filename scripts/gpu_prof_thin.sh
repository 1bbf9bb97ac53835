#!/bin/bash
# Round-2 diagnosis: --set full + source pages of the thin high-resolution layers (2..9 and 76..82).
mkdir -p gpurun_out
echo "=== layers 2..10 ==="
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel --launch-skip 81 --launch-count 9 \
   -o gpurun_out/r2_thin_a python scripts/one_forward.py 64 1 > gpurun_out/ncu_thin_a.log 2>&1
tail -2 gpurun_out/ncu_thin_a.log
echo "=== layers 76..82 ==="
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel --launch-skip 155 --launch-count 7 \
   -o gpurun_out/r2_thin_b python scripts/one_forward.py 64 1 > gpurun_out/ncu_thin_b.log 2>&1
tail -2 gpurun_out/ncu_thin_b.log
ls -la gpurun_out/*.ncu-rep
