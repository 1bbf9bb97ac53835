#!/bin/bash
# Round-2 evidence: --set full captures of every conv launch of one batch-64 forward + the mask kernel.
mkdir -p gpurun_out
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"conv_tc_kernel|conv1_tc|mask_" --launch-skip 82 --launch-count 82 \
   -o gpurun_out/r2_conv_full python scripts/one_forward.py 64 1 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
ncu -i gpurun_out/r2_conv_full.ncu-rep --page raw --csv > gpurun_out/r2_conv_full_raw.csv 2>/dev/null
python scripts/ncu_summary.py gpurun_out/r2_conv_full_raw.csv > gpurun_out/r2_conv_full_summary.txt
cat gpurun_out/r2_conv_full_summary.txt
ls -la gpurun_out/*.ncu-rep
