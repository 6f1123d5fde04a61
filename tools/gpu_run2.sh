#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
( time timeout 900 python -m pytest tests -m gpu -x -q --durations=8 ) > $OUT/pytest_gpu.log 2>&1
tail -15 $OUT/pytest_gpu.log
timeout 300 python tools/ab_inproc.py > $OUT/ab2.log 2> $OUT/ab2.err; cat $OUT/ab2.log; tail -3 $OUT/ab2.err
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu > $OUT/bench2.json 2> $OUT/bench2.err; tail -c 2500 $OUT/bench2.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 60 --csv --log-file $OUT/launches2.csv python tools/prof_one.py 80 > $OUT/prof_one.log 2>&1; tail -2 $OUT/prof_one.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_band -s 31 -c 2 -o $OUT/prof_band2 python tools/prof_one.py 40 > $OUT/ncu2.log 2>&1; tail -3 $OUT/ncu2.log
