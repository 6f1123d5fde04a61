#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
( time timeout 1200 python -m pytest tests -m gpu -x -q --durations=3 ) > $OUT/pytest_gpu2.log 2>&1
tail -8 $OUT/pytest_gpu2.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 20 --warmup 5 > $OUT/bench11_n2.json 2> $OUT/bench11_n2.err; tail -3 $OUT/bench11_n2.err
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/bench11_n2.json") if l.startswith('{')][-1])
print('N=2 value',round(d['value']),'ms/step',round(d['ms_per_step'],4),'e2e',round(d['e2e']['value']),round(d['e2e']['ms_per_call'],1),'launches',d['gpu_launches'])
print('  parity',d['parity']); print('  c4',d.get('c4')); print('  stages',d['config'].get('stages_ms_per_step_rank0'))
PY
python tools/measure_peaks.py
