#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -s 60 -c 30 --csv --log-file $OUT/launches8_c4.csv python tools/prof_c4.py 20 > $OUT/prof_c4.log 2>&1; tail -2 $OUT/prof_c4.log
