#!/bin/bash
# N-GPU check of the default bench (inference + e2e + train legs under torchrun) and the 2-GPU DP test. NG=2|4|8
NG=${NG:-2}
mkdir -p gpurun_out
nvidia-smi -L | head -8
nvidia-smi topo -m 2>/dev/null | head -12
echo "=== DP test (2 GPUs) ==="
timeout 600 python -m pytest tests/test_gpu_train.py -q -m gpu -s -k "parallel" 2>&1 | grep -E "passed|failed|skipped|Error|dp results" | cut -c1-400
echo "=== bench N=$NG ==="
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $NG --steps ${STEPS:-10} --warmup 3 --latency 0 --no-pipeline > gpurun_out/bench_n$NG.json 2> gpurun_out/bench_n$NG.err
tail -3 gpurun_out/bench_n$NG.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_n$NG.json').read().strip().splitlines()[-1])
d.pop('per_layer',None)
for k in ('value','n_gpus','ms_per_step','e2e','e2e_reference_layout','scaling','clocks','train'): print(k, d.get(k))
PY
