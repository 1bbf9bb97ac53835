#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"conv1_tc|mask_kernel" --launch-skip 2 --launch-count 2 \
   -o /tmp/r1_conv1 python scripts/one_forward.py 64 1 > gpurun_out/ncu_conv1.log 2>&1
tail -2 gpurun_out/ncu_conv1.log
ncu -i /tmp/r1_conv1.ncu-rep --page details > gpurun_out/r1_conv1_mask_details.txt 2>/dev/null
ncu -i /tmp/r1_conv1.ncu-rep --page source --csv --kernel-name regex:conv1 > gpurun_out/r1_conv1_source.csv 2>/dev/null
ncu -i /tmp/r1_conv1.ncu-rep --page source --csv --kernel-name regex:mask > gpurun_out/r1_mask_source.csv 2>/dev/null
ls -la gpurun_out/r1_conv1* gpurun_out/r1_mask*
