#!/bin/bash
# test + bench cycle.  TESTS="tests/x.py tests/y.py" selects test files (default: whole GPU suite)
mkdir -p gpurun_out
echo "=== tests ==="
timeout 1500 python -m pytest ${TESTS:-tests} -q -m gpu -x ${PYTEST_ARGS} 2>&1 | tail -${TAIL:-15} | tee gpurun_out/tests.log
echo "=== bench ==="
timeout 900 python bench.py --steps ${STEPS:-10} --warmup 3 ${BENCH_ARGS} > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -5 gpurun_out/bench.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
except Exception as e:
    print('no bench line', e); raise SystemExit
pl=d.pop('per_layer')
for k in ("value","ms_per_step","e2e","e2e_reference_layout","e2e_pipeline",'gpu_launches','latency_batch1','cpu_baseline','clocks','train','stress'): print(k, d.get(k))
r=d['roofline']; print('roofline', r['achieved'], r['frac'], r['kernel_ms_per_step'], r['network_ms_per_step'])
print([(x['kernel'][:12], round(x['ms']*1e3,1), round(x.get('frac',0) or 0,3)) for x in d['roofline_extra']])
print(' '.join('%s:%.3f/%.0f'%(k,v['ms'],v['tflops']) for k,v in pl.items()))
PY
