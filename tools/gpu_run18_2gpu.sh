#!/bin/bash
# Session 4, call 5 (gpurun --gpus 2, ~30 s): the N = 2 bench line with the reused label buffer in the sharded e2e leg.
OUT=gpurun_out; mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 70 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu --c4-steps 0 > $OUT/r18_bench_n2.json 2> $OUT/r18_bench_n2.err
tail -2 $OUT/r18_bench_n2.err; python - <<'PY'
import json
try:
    d = json.loads([l for l in open("gpurun_out/r18_bench_n2.json") if l.startswith('{')][-1])
    print('N=2 value', round(d['value']), 'us/step', round(d['ms_per_step'] * 1e3, 1), 'e2e', round(d['e2e']['value']), round(d['e2e']['ms_per_call'], 1), 'launches', d['gpu_launches'], d['parity'])
except Exception as e:
    print('FAILED', e)
PY
