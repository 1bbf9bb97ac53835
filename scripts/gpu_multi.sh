#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L
echo "=== DP test (2 GPUs) ==="
timeout 600 python -m pytest tests/test_gpu_train.py -q -m gpu -s -k "parallel" 2>&1 | grep -E "passed|failed|skipped|Error|dp results" | cut -c1-400
echo "=== inference bench N=2 ==="
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --latency 0 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
tail -3 gpurun_out/bench_n2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n2.json').read().strip().splitlines()[-1])
d.pop('per_layer',None)
for k in ('value','n_gpus','ms_per_step','e2e','scaling','clocks'): print(k, d[k])
PY
echo "=== reference arm N=2 (rank 0 only) ==="
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 2>/dev/null | tail -1 | cut -c1-400
echo "=== train bench N=2 ==="
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --workload train --gpus 2 --steps 3 --warmup 1 2>/dev/null | tail -1 | cut -c1-600
