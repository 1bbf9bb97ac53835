#!/bin/bash
# Round-2 evidence: GPU test suite, smoke, the driver's bench command, launch list of the same command, planner A/B,
# --set full captures (details + SASS/source pages) of representative conv launches and the mask kernel.
mkdir -p gpurun_out
echo "=== tests ==="
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -3 | tee gpurun_out/r2_tests.log
echo "=== smoke ==="
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
echo "=== bench (the driver's command) ==="
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_final_bench.json 2> gpurun_out/bench.err
tail -3 gpurun_out/bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_final_bench.json').read().strip().splitlines()[-1])
pl=d.pop('per_layer')
for k in ('value','ms_per_step','clocks','e2e','e2e_reference_layout','e2e_pipeline','gpu_launches','latency_batch1','cpu_baseline','train','stress'): print(k, d[k])
print('roofline', d['roofline'])
print([(x['kernel'][:12], round(x['ms']*1e3,1), round(x.get('frac',0) or 0,3)) for x in d['roofline_extra']])
print(' '.join('%s:%.3f/%.0f'%(k,v['ms'],v['tflops']) for k,v in pl.items()))
PY
echo "=== inference launch list (2 timed steps of the bench command; 85 launches per forward) ==="
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 340 -c 170 --csv \
   --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 3 --no-steady --latency 0 --no-cpu --no-pipeline --no-train --no-stress > gpurun_out/ncu_bench.log 2>&1
wc -l gpurun_out/r2_launches.csv
python scripts/launch_agg.py gpurun_out/r2_launches.csv 2>/dev/null | head -12
echo "=== planner A/B ==="
timeout 400 python scripts/ab_opts.py auto: nocta2:tc_cta2=0 nosplit:tc_split_n=0 nohalo:tc_halo=0 > gpurun_out/r2_ab_cta2.txt 2>&1
tail -2 gpurun_out/r2_ab_cta2.txt
echo "=== --set full: layers 12 (CTA pair + halo + split-N), 58, 81 (fused tail), 4 ==="
LAYERS="12 58 81 4" TOP=25 bash scripts/gpu_prof_one.sh > gpurun_out/r2_prof_layers.txt 2>&1
grep -E "layer|Duration|Grid Size|Registers" gpurun_out/r2_prof_layers.txt | head -24
for L in 12; do grep -c "UTCHMMA.2CTA" gpurun_out/r2_conv_L${L}_source.csv; done
echo "=== --set full: mask kernel ==="
timeout 300 ncu --set full --clock-control none -k regex:mask_ --launch-skip 1 --launch-count 1 -f -o /tmp/r2_mask python scripts/one_forward.py 64 1 > /tmp/ncu_mask.log 2>&1
ncu -i /tmp/r2_mask.ncu-rep --page details > gpurun_out/r2_mask_details.txt 2>/dev/null
grep -E "Duration|DRAM Throughput|Memory Throughput" gpurun_out/r2_mask_details.txt | head -5
ls -la gpurun_out | head -40
