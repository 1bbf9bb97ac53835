#!/bin/bash
# Round-2 evidence: launch list of the bench command's timed steps and --set full captures of the dominant kernels.
mkdir -p gpurun_out
echo "=== inference launch list (bench.py --steps 2 --warmup 3: 3 warm-up + 2 timed forwards, 85 launches each) ==="
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 255 -c 170 --csv \
   --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 3 --latency 0 --no-cpu --no-pipeline --no-train --no-stress > gpurun_out/ncu_bench.log 2>&1
wc -l gpurun_out/r2_launches.csv
echo "=== set full: conv_tc layers 2..9 and 76..81 + mask kernel (second forward of one_forward.py) ==="
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"conv_tc_kernel|conv1_tc|mask_kernel" --launch-skip 81 --launch-count 82 \
   -o gpurun_out/r2_conv_full python scripts/one_forward.py 64 1 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
ncu -i gpurun_out/r2_conv_full.ncu-rep --page raw --csv > gpurun_out/r2_conv_full_raw.csv 2>/dev/null
python scripts/ncu_summary.py gpurun_out/r2_conv_full_raw.csv > gpurun_out/r2_conv_full_summary.txt
tail -5 gpurun_out/r2_conv_full_summary.txt
ls -la gpurun_out/*.ncu-rep
