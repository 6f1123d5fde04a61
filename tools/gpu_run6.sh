#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
( time timeout 900 python -m pytest tests -m gpu -x -q --durations=3 ) > $OUT/pytest_gpu.log 2>&1
tail -8 $OUT/pytest_gpu.log
( time timeout 600 python bench.py --steps 20 --warmup 5 ) > $OUT/bench6.json 2> $OUT/bench6.err; tail -3 $OUT/bench6.err; python -c "
import json
d=json.loads([l for l in open('$OUT/bench6.json') if l.startswith('{')][-1])
print('value',d['value'],'ms/step',d['ms_per_step'],'e2e',d['e2e']['value'],d['e2e']['ms_per_call'],'pinned',d.get('e2e_pinned'))
print('parity',d['parity']); print('c4',d.get('c4')); print('stages',d['roofline']['stages_ms_per_step']); print('cpu',d.get('cpu_baseline'))"
nvidia-smi --query-gpu=memory.used --format=csv,noheader
