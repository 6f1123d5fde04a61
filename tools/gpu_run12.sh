#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 20 --warmup 5 > $OUT/bench12_n2.json 2> $OUT/bench12_n2.err; tail -3 $OUT/bench12_n2.err
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/bench12_n2.json") if l.startswith('{')][-1])
print('N=2 value',round(d['value']),'e2e',round(d['e2e']['value']),round(d['e2e']['ms_per_call'],1)); print('  parity',d['parity']); print('  c4',d.get('c4'))
PY
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_lloyd.py -m gpu -x -q -k "two_gpus or two_band" 2>&1 | tail -3
