#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
( time timeout 900 python -m pytest tests -m gpu -x -q --durations=3 ) > $OUT/pytest_gpu.log 2>&1
tail -8 $OUT/pytest_gpu.log
timeout 300 python tools/prof_c4.py 20 2>&1 | tail -1
timeout 600 python tools/bench_batch.py --meshes 32 --iters 100 > $OUT/batch9.json 2> $OUT/batch9.err; tail -2 $OUT/batch9.err; python -c "
import json;d=json.loads(open('$OUT/batch9.json').read().strip().splitlines()[-1]);d.pop('sites_sha1_per_mesh');print(d)"
