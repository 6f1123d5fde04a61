#!/bin/bash
# Session 3, call 1: the whole GPU suite with the band order on, A/B of the band order, the official bench line, the
# pending A/B measurements of the streaming kernels / JFA family / host pipeline, launch list + one full capture.
OUT=gpurun_out; mkdir -p $OUT
export PYTHONUNBUFFERED=1
T0=$(date +%s)
( time SRM_BAND_ORDER=1 timeout 600 python -m pytest tests -m gpu -q --durations=5 ) > $OUT/r13_pytest_order1.log 2>&1
tail -12 $OUT/r13_pytest_order1.log; echo "== t=$(( $(date +%s) - T0 ))s"
timeout 240 python tools/bench_band_order.py > $OUT/r13_band_order.json 2> $OUT/r13_band_order.err; cat $OUT/r13_band_order.err | tail -6; echo "== t=$(( $(date +%s) - T0 ))s"
timeout 420 python bench.py --gpus 1 --steps 20 --warmup 5 > $OUT/r13_bench_n1.json 2> $OUT/r13_bench_n1.err; tail -2 $OUT/r13_bench_n1.err
SRM_BAND_ORDER=1 timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu > $OUT/r13_bench_n1_order1.json 2> $OUT/r13_bench_n1_order1.err; tail -2 $OUT/r13_bench_n1_order1.err
python - <<'PY'
import json
for f in ("r13_bench_n1", "r13_bench_n1_order1"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/{f}.json") if l.startswith('{')][-1])
        print(f, 'value', round(d['value']), 'us/step', round(d['ms_per_step'] * 1e3, 1), 'e2e', round(d['e2e']['value']), round(d['e2e']['ms_per_call'], 1),
              'launches', d['gpu_launches'], 'clocks', d['clocks'])
        print('   parity', d['parity']); print('   c4', {k: d['c4'].get(k) for k in ('value', 'ms_per_step', 'k_band_ms_per_rank', 'error')} if d.get('c4') else None)
        print('   roofline', {k: d['roofline'].get(k) for k in ('achieved', 'frac', 'ms_per_launch')}, d.get('cpu_baseline', {}).get('value'))
    except Exception as e:
        print(f, 'FAILED', e)
PY
echo "== t=$(( $(date +%s) - T0 ))s"
timeout 200 python tools/bench_streams.py > $OUT/r13_streams.json 2> $OUT/r13_streams.err; cat $OUT/r13_streams.json | head -c 1500; echo
timeout 240 python tools/bench_jfa.py > $OUT/r13_jfa.json 2> $OUT/r13_jfa.err; python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r13_jfa.json").read().strip().splitlines()[-1])
    for m in ("mode0", "mode1"):
        print(m, round(d[m]["total_ms"], 3), "ms", d[m]["launches"], "launches", [(x["launch"], x["ms"], x["frac_of_peak"]) for x in d[m]["per_launch"]])
    print({k: d[k] for k in d if k not in ("mode0", "mode1", "schedule")})
except Exception as e:
    print("jfa FAILED", e)
PY
echo "== t=$(( $(date +%s) - T0 ))s"
timeout 240 python tools/bench_host.py > $OUT/r13_host.json 2> $OUT/r13_host.err; tail -9 $OUT/r13_host.err; echo "== t=$(( $(date +%s) - T0 ))s"
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/r13_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --c4-steps 0 > $OUT/r13_ncu_list.log 2>&1; tail -1 $OUT/r13_ncu_list.log | head -c 300; echo; echo "== t=$(( $(date +%s) - T0 ))s"
SRM_BAND_ORDER=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:^k_band -s 31 -c 3 -o $OUT/r13_prof_band python tools/prof_one.py 40 > $OUT/r13_ncu_band.log 2>&1; tail -2 $OUT/r13_ncu_band.log; echo "== t=$(( $(date +%s) - T0 ))s"
