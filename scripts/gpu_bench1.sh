#!/bin/bash
mkdir -p gpurun_out
echo "=== bench ==="
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench1.json 2> gpurun_out/bench1.err
tail -5 gpurun_out/bench1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench1.json').read().strip().splitlines()[-1])
pl=d.pop('per_layer')
print(json.dumps(d, indent=1))
print(' '.join('%s:%.3f/%.0f'%(k,v['ms'],v['tflops']) for k,v in pl.items()))
PY
echo "=== ncu launch list ==="
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 261 -c 100 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 1 --warmup 3 --latency 0 --no-cpu > gpurun_out/ncu_bench.log 2>&1
tail -3 gpurun_out/ncu_bench.log
wc -l gpurun_out/launches_r1.csv
