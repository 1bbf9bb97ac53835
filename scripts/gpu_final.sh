#!/bin/bash
# Round-end evidence: full GPU test suite, smoke, default bench lines, refreshed launch lists.
mkdir -p gpurun_out
echo "=== tests ==="
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -3 | tee gpurun_out/tests.log
echo "=== smoke ==="
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
echo "=== bench (default flags) ==="
timeout 900 python bench.py > gpurun_out/r1_final_bench.json 2> gpurun_out/bench.err
tail -3 gpurun_out/bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r1_final_bench.json').read().strip().splitlines()[-1])
pl=d.pop('per_layer')
for k in ('value','ms_per_step','clocks','e2e','e2e_pipeline','gpu_launches','latency_batch1','cpu_baseline'): print(k, d[k])
r=d['roofline']; print('roofline', r['achieved'], r['frac'], r['kernel_ms_per_step'], r['network_ms_per_step'])
print([(x['kernel'][:12], round(x['ms']*1e3,1), round(x.get('frac',0) or 0,3)) for x in d['roofline_extra']])
print(' '.join('%s:%.3f/%.0f'%(k,v['ms'],v['tflops']) for k,v in pl.items()))
PY
echo "=== bench train ==="
timeout 600 python bench.py --workload train --steps 20 --warmup 3 2>/dev/null | tail -1 > gpurun_out/r1_train_bench.json; cut -c60-330 gpurun_out/r1_train_bench.json
echo "=== inference launch list ==="
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 258 -c 172 --csv \
   --log-file gpurun_out/r1_launches.csv python bench.py --steps 2 --warmup 3 --latency 0 --no-cpu --no-pipeline > gpurun_out/ncu_bench.log 2>&1
wc -l gpurun_out/r1_launches.csv
echo "=== training launch list ==="
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
   --log-file gpurun_out/r1_train_launches.csv python bench.py --workload train --steps 1 --warmup 3 > gpurun_out/ncu_train.log 2>&1
wc -l gpurun_out/r1_train_launches.csv
