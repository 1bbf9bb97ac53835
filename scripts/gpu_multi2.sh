#!/bin/bash
# N=2 checks: data-parallel training on the tensor-core engine, sharded inference, the 2-GPU DP test
mkdir -p gpurun_out
nvidia-smi --query-gpu=name --format=csv,noheader | head -4
echo "=== DP test (fp32 engine) ==="
timeout 600 python -m pytest tests/test_gpu_train.py -q -m gpu -k two_gpus 2>&1 | tail -2
echo "=== train N=2 bf16 ==="
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --workload train --steps 10 --warmup 3 2>&1 | tail -1 | cut -c1-700 | tee gpurun_out/train_n2.json
echo "=== inference N=2 ==="
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 --latency 0 2> gpurun_out/bench_n2.err | tail -1 > gpurun_out/bench_n2.json
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n2.json').read().strip().splitlines()[-1]); d.pop('per_layer',None)
print({k:d[k] for k in ('value','n_gpus','ms_per_step','e2e','clocks')})
PY
