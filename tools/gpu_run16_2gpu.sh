#!/bin/bash
# Session 4, call 3 (gpurun --gpus 2): the multi-process parity tests against the single-process oracle and the bench line
# at N = 2 (8192^2 + the C4 leg) on the shipped build.
OUT=gpurun_out; mkdir -p $OUT
export PYTHONUNBUFFERED=1
T0=$(date +%s)
( time timeout 170 python -m pytest tests/test_gpu_multi.py -m gpu -q --durations=3 ) > $OUT/r16_pytest_multi.log 2>&1
tail -8 $OUT/r16_pytest_multi.log; echo "== t=$(( $(date +%s) - T0 ))s"
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu > $OUT/r16_bench_n2.json 2> $OUT/r16_bench_n2.err
tail -3 $OUT/r16_bench_n2.err; python - <<'PY'
import json
try:
    d = json.loads([l for l in open("gpurun_out/r16_bench_n2.json") if l.startswith('{')][-1])
    print('N=2 value', round(d['value']), 'us/step', round(d['ms_per_step'] * 1e3, 1), 'e2e', round(d['e2e']['value']), round(d['e2e']['ms_per_call'], 1), 'launches', d['gpu_launches'])
    print('   parity', d['parity']); print('   c4', {k: d['c4'].get(k) for k in ('value', 'ms_per_step', 'k_band_ms_per_rank', 'bands', 'error', 'sites_sha1_after')} if d.get('c4') else None)
except Exception as e:
    print('FAILED', e)
PY
echo "== t=$(( $(date +%s) - T0 ))s"
