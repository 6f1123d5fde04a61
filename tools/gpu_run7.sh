#!/bin/bash
# 2-GPU session: full GPU suite (incl. the 2-GPU parity tests), bench at N=1 and N=2
OUT=gpurun_out; mkdir -p $OUT
( time timeout 1200 python -m pytest tests -m gpu -x -q --durations=5 ) > $OUT/pytest_gpu2.log 2>&1
tail -12 $OUT/pytest_gpu2.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu > $OUT/bench7_n1.json 2> $OUT/bench7_n1.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > $OUT/bench7_n2.json 2> $OUT/bench7_n2.err; tail -5 $OUT/bench7_n2.err
python - <<'PY'
import json
for f in ("bench7_n1","bench7_n2"):
    try:
        d=json.loads([l for l in open(f"gpurun_out/{f}.json") if l.startswith('{')][-1])
        print(f, 'value',round(d['value']),'ms/step',round(d['ms_per_step'],4),'e2e',round(d['e2e']['value']),round(d['e2e']['ms_per_call'],1),'launches',d['gpu_launches'])
        print('  parity',d['parity']); print('  c4',d.get('c4')); print('  stages',d['roofline']['stages_ms_per_step'] if 'roofline' in d else None)
    except Exception as e: print(f, 'ERR', e)
PY
