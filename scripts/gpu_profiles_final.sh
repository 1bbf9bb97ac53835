#!/bin/bash
# Round-1 evidence.  Launch lists of the bench commands (inference: 2 timed steps; training: all launches of
# `--steps 1 --warmup 3`), and --set full captures of the dominant kernels.
mkdir -p gpurun_out
echo "=== inference launch list (bench.py --steps 2 --warmup 3, timed region) ==="
# 3 warm-up forwards = 3 x 86 launches are skipped, the 2 timed steps (172 launches) are listed
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 258 -c 172 --csv \
   --log-file gpurun_out/r1_launches.csv python bench.py --steps 2 --warmup 3 --latency 0 --no-cpu > gpurun_out/ncu_bench.log 2>&1
wc -l gpurun_out/r1_launches.csv
echo "=== training launch list (bench.py --workload train --steps 1 --warmup 3) ==="
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
   --log-file gpurun_out/r1_train_launches.csv python bench.py --workload train --steps 1 --warmup 3 > gpurun_out/ncu_train.log 2>&1
wc -l gpurun_out/r1_train_launches.csv
echo "=== set full: wgrad (one training step: 30 launches) ==="
timeout 900 ncu --set full --clock-control none --import-source on -k regex:wgrad_tc_kernel --launch-skip 90 --launch-count 30 \
   -o /tmp/r1_wgrad_full python bench.py --workload train --steps 1 --warmup 3 > gpurun_out/ncu_wgrad.log 2>&1
ncu -i /tmp/r1_wgrad_full.ncu-rep --page raw --csv > gpurun_out/r1_wgrad_full_raw.csv 2>/dev/null
ncu -i /tmp/r1_wgrad_full.ncu-rep --page details > gpurun_out/r1_wgrad_full_details.txt 2>/dev/null
ncu -i /tmp/r1_wgrad_full.ncu-rep --page source --csv --kernel-id :::25 > gpurun_out/r1_wgrad_L58_source.csv 2>/dev/null
echo "=== set full: conv_tc layers 54..82 + mask kernel (second forward) ==="
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"conv_tc_kernel|conv1_tc|mask_kernel" --launch-skip 136 --launch-count 30 \
   -o /tmp/r1_conv_full python scripts/one_forward.py 64 1 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
ncu -i /tmp/r1_conv_full.ncu-rep --page raw --csv > gpurun_out/r1_conv_full_raw.csv 2>/dev/null
ncu -i /tmp/r1_conv_full.ncu-rep --page details > gpurun_out/r1_conv_full_details.txt 2>/dev/null
echo "=== set full: conv1 ==="
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv1_tc --launch-skip 1 --launch-count 1 \
   -o /tmp/r1_conv1_full python scripts/one_forward.py 64 1 > /dev/null 2>&1
ncu -i /tmp/r1_conv1_full.ncu-rep --page details > gpurun_out/r1_conv1_full_details.txt 2>/dev/null
ls -la gpurun_out/ | tail -20
