#!/bin/bash
# Session 4, call 1: the whole GPU suite on the new defaults (k_prefix / k_expand builds, block-wise host scans, dense
# centroid pass), the official bench line, where the time of a gCVT call goes (SRM_TRACE), the streaming kernels incl. the
# centroid pass, 16-row bands A/B, launch list of the loop + one full capture of the streaming kernels.
OUT=gpurun_out; mkdir -p $OUT
export PYTHONUNBUFFERED=1
T0=$(date +%s)
( time timeout 420 python -m pytest tests -m gpu -q --durations=5 ) > $OUT/r14_pytest.log 2>&1
tail -14 $OUT/r14_pytest.log; echo "== t=$(( $(date +%s) - T0 ))s"
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 > $OUT/r14_bench_n1.json 2> $OUT/r14_bench_n1.err; tail -2 $OUT/r14_bench_n1.err
python - <<'PY'
import json
for f in ("r14_bench_n1",):
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
SRM_TRACE=1 timeout 120 python tools/bench_host.py --quick > $OUT/r14_host.json 2> $OUT/r14_host.err; grep -v "^\[bench\]" $OUT/r14_host.err | tail -40; echo "== t=$(( $(date +%s) - T0 ))s"
timeout 120 python tools/bench_streams.py > $OUT/r14_streams.json 2> $OUT/r14_streams.err; cat $OUT/r14_streams.json | head -c 2500; echo
SRM_CEN_WAVES=4 timeout 120 python tools/bench_streams.py > $OUT/r14_streams_w4.json 2> $OUT/r14_streams_w4.err; python -c "
import json; d=json.load(open('gpurun_out/r14_streams_w4.json')); print('waves=4', d.get('centroid_dense'))"
echo "== t=$(( $(date +%s) - T0 ))s"
timeout 90 python tools/bench_band_order.py --orders 0 --steps 60 > $OUT/r14_rpw1.json 2> $OUT/r14_rpw1.err; tail -1 $OUT/r14_rpw1.err
SRM_BAND_RPW=2 timeout 90 python tools/bench_band_order.py --orders 0 --steps 60 > $OUT/r14_rpw2.json 2> $OUT/r14_rpw2.err; tail -1 $OUT/r14_rpw2.err
echo "== t=$(( $(date +%s) - T0 ))s"
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'^k_' -c 400 --csv --log-file $OUT/r14_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --c4-steps 0 > $OUT/r14_ncu_list.log 2>&1; tail -1 $OUT/r14_ncu_list.log | head -c 300; echo; echo "== t=$(( $(date +%s) - T0 ))s"
timeout 150 ncu --set full --clock-control none --import-source on -k regex:'^k_(prefix|expand|centroid)' -c 6 -o $OUT/r14_prof_streams python tools/prof_streams.py > $OUT/r14_ncu_streams.log 2>&1; tail -2 $OUT/r14_ncu_streams.log; echo "== t=$(( $(date +%s) - T0 ))s"
