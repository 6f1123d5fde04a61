#!/bin/bash
# round-2 GPU session 1: full GPU test suite, kernel variants A/B, ablation, bench (both arms)
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader > $OUT/gpu.txt; free -g | head -2 >> $OUT/gpu.txt; nproc >> $OUT/gpu.txt
( time timeout 900 python -m pytest tests -m gpu -x -q --durations=15 ) > $OUT/pytest_gpu.log 2>&1
tail -30 $OUT/pytest_gpu.log
timeout 300 python tools/ab_inproc.py nst1 nst2 cta3 > $OUT/ab1.log 2> $OUT/ab1.err; cat $OUT/ab1.log
timeout 200 python tools/ablate_band.py > $OUT/ablate1.log 2> $OUT/ablate1.err; cat $OUT/ablate1.log
timeout 300 python bench.py --steps 20 --warmup 5 > $OUT/bench1.json 2> $OUT/bench1.err; tail -c 3000 $OUT/bench1.json
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > $OUT/bench1_ref.json 2> $OUT/bench1_ref.err; tail -c 1500 $OUT/bench1_ref.json
