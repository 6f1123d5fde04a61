#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
( time timeout 900 python -m pytest tests -m gpu -x -q --durations=3 ) > $OUT/pytest_gpu.log 2>&1
tail -8 $OUT/pytest_gpu.log
SRM_TRACE=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu > $OUT/bench5.json 2> $OUT/bench5.err; grep srm_gcvt $OUT/bench5.err | tail -24; python -c "
import json;d=json.load(open('$OUT/bench5.json'));print('20 steps:',d['value'],d['ms_per_step'],d['e2e'],d.get('e2e_pinned'))"
for T in 4 16; do SRM_HOST_THREADS=$T timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('threads $T:',d['e2e']['value'],d['e2e']['ms_per_call'])"; done
