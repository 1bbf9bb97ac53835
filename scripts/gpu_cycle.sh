#!/bin/bash
# test + bench cycle
mkdir -p gpurun_out
echo "=== tests ==="
timeout 1200 python -m pytest tests -q -m gpu -x 2>&1 | tail -8 | tee gpurun_out/tests.log
echo "=== bench ==="
timeout 900 python bench.py --steps 5 --warmup 3 ${BENCH_ARGS} > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -5 gpurun_out/bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
pl=d.pop('per_layer')
for k in ("value","ms_per_step","e2e","e2e_pipeline",'gpu_launches','latency_batch1','cpu_baseline','clocks'): print(k, d[k])
r=d['roofline']; print('roofline', r['achieved'], r['frac'], r['kernel_ms_per_step'], r['network_ms_per_step'])
print(d['roofline_extra'])
print(' '.join('%s:%.3f/%.0f'%(k,v['ms'],v['tflops']) for k,v in pl.items()))
PY
