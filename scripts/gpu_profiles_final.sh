#!/bin/bash
# Round-1 evidence: launch list of the bench command, one --set full capture of the conv kernel.
mkdir -p gpurun_out
echo "=== launch list (bench.py --steps 2 --warmup 3, timed region) ==="
# warm-up forwards: 3 (value) ; skip their 86*3 launches, list the 2 timed steps
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 258 -c 172 --csv \
   --log-file gpurun_out/r1_launches.csv python bench.py --steps 2 --warmup 3 --latency 0 --no-cpu > gpurun_out/ncu_bench.log 2>&1
wc -l gpurun_out/r1_launches.csv
echo "=== set full: conv_tc launches of layers 54..81 (second forward) ==="
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel --launch-skip 133 --launch-count 28 \
   -o /tmp/r1_conv_full python scripts/one_forward.py 64 1 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
ls -la /tmp/r1_conv_full.ncu-rep
ncu -i /tmp/r1_conv_full.ncu-rep --page raw --csv > gpurun_out/r1_conv_full_raw.csv 2>/dev/null
ncu -i /tmp/r1_conv_full.ncu-rep --page details > gpurun_out/r1_conv_full_details.txt 2>/dev/null
ncu -i /tmp/r1_conv_full.ncu-rep --page source --csv --kernel-id :::1 > gpurun_out/r1_conv_L54_source.csv 2>/dev/null
ls -la gpurun_out/
