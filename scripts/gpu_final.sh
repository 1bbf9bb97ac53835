#!/bin/bash
# full GPU test suite + the default bench lines (inference, training, reference arm) + smoke
mkdir -p gpurun_out
echo "=== tests ==="
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -4 | tee gpurun_out/tests.log
echo "=== smoke ==="
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "=== bench (default flags) ==="
timeout 900 python bench.py > gpurun_out/r1_final_bench.json 2> gpurun_out/bench.err
tail -3 gpurun_out/bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r1_final_bench.json').read().strip().splitlines()[-1])
d.pop('per_layer')
print(json.dumps(d)[:3000])
PY
echo "=== bench train ==="
timeout 600 python bench.py --workload train --steps 20 --warmup 3 2>/dev/null | tail -1 > gpurun_out/r1_train_bench.json; cut -c1-900 gpurun_out/r1_train_bench.json
echo "=== reference arm ==="
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 | cut -c1-600
