#!/bin/bash
# Session 4, call 2: the whole GPU suite on the shipped defaults (k_expand3, k_centroid_dense2, k_prefix<1>), the official
# bench line, the streaming kernels in all builds, launch list of the bench + one full capture of the streaming kernels.
OUT=gpurun_out; mkdir -p $OUT
export PYTHONUNBUFFERED=1
T0=$(date +%s)
( time timeout 420 python -m pytest tests -m gpu -q --durations=5 ) > $OUT/r15_pytest.log 2>&1
tail -14 $OUT/r15_pytest.log; grep -E "FAILED|ERROR" $OUT/r15_pytest.log | head -20; echo "== t=$(( $(date +%s) - T0 ))s"
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 > $OUT/r15_bench_n1.json 2> $OUT/r15_bench_n1.err; tail -2 $OUT/r15_bench_n1.err
python - <<'PY'
import json
for f in ("r15_bench_n1",):
    try:
        d = json.loads([l for l in open(f"gpurun_out/{f}.json") if l.startswith('{')][-1])
        print(f, 'value', round(d['value']), 'us/step', round(d['ms_per_step'] * 1e3, 1), 'e2e', round(d['e2e']['value']), round(d['e2e']['ms_per_call'], 1),
              'pinned', d.get('e2e_pinned'), 'launches', d['gpu_launches'], 'clocks', d['clocks'])
        print('   parity', d['parity']); print('   c4', {k: d['c4'].get(k) for k in ('value', 'ms_per_step', 'k_band_ms_per_rank', 'error')} if d.get('c4') else None)
        print('   roofline', {k: d['roofline'].get(k) for k in ('achieved', 'frac', 'ms_per_launch')}, d.get('cpu_baseline', {}).get('value'))
    except Exception as e:
        print(f, 'FAILED', e)
PY
echo "== t=$(( $(date +%s) - T0 ))s"
timeout 120 python tools/bench_streams.py > $OUT/r15_streams.json 2> $OUT/r15_streams.err; cat $OUT/r15_streams.json | head -c 3500; echo; tail -3 $OUT/r15_streams.err
for w in 2 8; do SRM_CEN_WAVES=$w timeout 120 python tools/bench_streams.py > $OUT/r15_streams_w$w.json 2> $OUT/r15_streams_w$w.err; python -c "
import json; d=json.load(open('gpurun_out/r15_streams_w$w.json')); print('waves=$w', d.get('centroid_dense_v0'), d.get('centroid_dense_v1'))"; done
echo "== t=$(( $(date +%s) - T0 ))s"
timeout 150 ncu --set full --clock-control none --import-source on -k regex:'^k_(prefix|expand|centroid)' -c 6 -o $OUT/r15_prof_streams python tools/prof_streams.py > $OUT/r15_ncu_streams.log 2>&1; tail -2 $OUT/r15_ncu_streams.log; echo "== t=$(( $(date +%s) - T0 ))s"
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'^k_' -c 400 --csv --log-file $OUT/r15_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --c4-steps 0 > $OUT/r15_ncu_list.log 2>&1; tail -1 $OUT/r15_ncu_list.log | head -c 200; echo; echo "== t=$(( $(date +%s) - T0 ))s"
