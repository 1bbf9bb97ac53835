#!/bin/bash
# --set full + source-page hot spots of selected conv_tc launches of the second forward.
# usage: LAYERS="81 4 2" bash scripts/gpu_prof_one.sh     (layer n = conv_tc launch n-2 of a forward; 80 per forward)
mkdir -p gpurun_out
for L in ${LAYERS:-81}; do
  SKIP=$((80 + L - 2))
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel --launch-skip $SKIP --launch-count 1 \
     -f -o /tmp/r2_L$L python scripts/one_forward.py 64 1 > /tmp/ncu_L$L.log 2>&1
  tail -1 /tmp/ncu_L$L.log
  ncu -i /tmp/r2_L$L.ncu-rep --page source --csv > gpurun_out/r2_conv_L${L}_source.csv 2>/dev/null
  ncu -i /tmp/r2_L$L.ncu-rep --page details > gpurun_out/r2_conv_L${L}_details.txt 2>/dev/null
  echo "=== layer $L hot instructions ==="
  python scripts/ncu_hot.py gpurun_out/r2_conv_L${L}_source.csv ${TOP:-45}
  grep -E "Duration|Executed Ipc Active|Warp Cycles Per Issued|No Eligible|Registers Per Thread|Dynamic Shared Memory Per Block|Grid Size" gpurun_out/r2_conv_L${L}_details.txt | head -12
done
