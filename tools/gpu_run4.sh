#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 600 ncu --set full --clock-control none --import-source on -k regex:^k_ -s 155 -c 5 -o $OUT/prof_step4 python tools/prof_one.py 40 > $OUT/ncu4.log 2>&1; tail -3 $OUT/ncu4.log
