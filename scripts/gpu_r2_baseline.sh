#!/bin/bash
# Round-2 checkpoint: GPU test suite, smoke, default bench line, inference launch list.
mkdir -p gpurun_out
echo "=== tests ==="
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -8 | tee gpurun_out/tests.log
echo "=== smoke ==="
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "=== bench (default flags) ==="
timeout 900 python bench.py > gpurun_out/r2_bench.json 2> gpurun_out/bench.err
tail -3 gpurun_out/bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench.json').read().strip().splitlines()[-1])
pl=d.pop('per_layer')
for k in ("value","ms_per_step","e2e","e2e_reference_layout","e2e_pipeline",'gpu_launches','latency_batch1','cpu_baseline','clocks','train','stress'): print(k, d.get(k))
r=d['roofline']; print('roofline', r)
print([(x['kernel'][:12], round(x['ms']*1e3,1), round(x.get('frac',0) or 0,3)) for x in d['roofline_extra']])
print(' '.join('%s:%.3f/%.0f'%(k,v['ms'],v['tflops']) for k,v in pl.items()))
PY
echo "=== inference launch list ==="
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 255 -c 170 --csv \
   --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 3 --latency 0 --no-cpu --no-pipeline --no-train --no-stress > gpurun_out/ncu_bench.log 2>&1
wc -l gpurun_out/r2_launches.csv
