mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv1_tc --launch-skip 1 --launch-count 1 -f -o /tmp/r2_c1 python scripts/one_forward.py 64 1 > /tmp/ncu_c1.log 2>&1
ncu -i /tmp/r2_c1.ncu-rep --page source --csv > gpurun_out/r2_conv1_source.csv 2>/dev/null
ncu -i /tmp/r2_c1.ncu-rep --page details > gpurun_out/r2_conv1_details.txt 2>/dev/null
grep -E "Duration|Executed Ipc|Registers|Issue Slots Busy|No Eligible" gpurun_out/r2_conv1_details.txt | head
