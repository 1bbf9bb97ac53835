#!/bin/bash
mkdir -p gpurun_out
timeout 1500 ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section WarpStateStats --section LaunchStats --section Occupancy --section SchedulerStats \
   --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts.sum,smsp__inst_executed.sum \
   --clock-control none -k regex:'conv|mask|decode|nms|finalize' --launch-skip 86 --launch-count 86 \
   -o /tmp/r1_sel python scripts/one_forward.py 64 1 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
ncu -i /tmp/r1_sel.ncu-rep --page raw --csv > gpurun_out/r1_sel_raw.csv 2>/dev/null
ls -la gpurun_out/r1_sel_raw.csv
