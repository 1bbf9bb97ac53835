#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
   -k regex:"p1_|bn_|wgrad" --log-file gpurun_out/r1_train_ew.csv python bench.py --workload train --steps 1 --warmup 3 > gpurun_out/ncu_train.log 2>&1
tail -1 gpurun_out/ncu_train.log | cut -c1-200
wc -l gpurun_out/r1_train_ew.csv
