#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
( time timeout 900 python -m pytest tests -m gpu -x -q --durations=5 ) > $OUT/pytest_gpu.log 2>&1
tail -12 $OUT/pytest_gpu.log
timeout 300 python tools/ab_inproc.py nopdl > $OUT/ab3.log 2> $OUT/ab3.err; cat $OUT/ab3.log; tail -3 $OUT/ab3.err
timeout 200 python tools/ablate_band.py > $OUT/ablate3.log 2> $OUT/ablate3.err; cat $OUT/ablate3.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu > $OUT/bench3.json 2> $OUT/bench3.err; tail -c 1800 $OUT/bench3.json
timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu > $OUT/bench3_200.json 2> $OUT/bench3_200.err; python -c "
import json;d=json.load(open('$OUT/bench3_200.json'));print('200 steps:',d['value'],d['ms_per_step'],d['e2e'],d.get('e2e_pinned'),d['roofline']['stages_ms_per_step'])"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -s 200 -c 60 --csv --log-file $OUT/launches3.csv python tools/prof_one.py 60 > $OUT/prof_one.log 2>&1; tail -2 $OUT/prof_one.log
