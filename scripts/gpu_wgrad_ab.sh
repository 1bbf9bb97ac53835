#!/bin/bash
mkdir -p gpurun_out
for o in 1 2; do
DY_OPTS="wgrad_fuse_kw=$o" timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv -k regex:wgrad --launch-skip 90 --launch-count 30 \
   --log-file gpurun_out/wgrad_fuse$o.csv python bench.py --workload train --steps 1 --warmup 3 > /dev/null 2>&1
python - <<PY
import csv
rows=list(csv.reader(open('gpurun_out/wgrad_fuse$o.csv')))
hi=[i for i,r in enumerate(rows) if 'Kernel Name' in r][0]
hdr=rows[hi]; vi=hdr.index('Metric Value'); ui=hdr.index('Metric Unit')
us=[float(r[vi].replace(',',''))*{'ns':1e-3,'us':1,'ms':1e3}.get(r[ui],1) for r in rows[hi+1:] if len(r)>vi]
print('fuse=$o  total %.0f us :'%sum(us), ' '.join('%.0f'%u for u in us))
PY
done
DY_OPTS="wgrad_fuse_kw=1" timeout 300 python bench.py --workload train --steps 30 --warmup 5 2>&1 | tail -1 | cut -c70-200
DY_OPTS="wgrad_fuse_kw=2" timeout 300 python bench.py --workload train --steps 30 --warmup 5 2>&1 | tail -1 | cut -c70-200
