#!/bin/bash
# A/B of kernel variants (build/variants/libsrm_*.so, selected with SRM_LIB): parity subset + bench for each.
OUT=gpurun_out; mkdir -p $OUT
for v in base "$@"; do
  if [ $v = base ]; then unset SRM_LIB; else export SRM_LIB=$PWD/build/variants/libsrm_$v.so; fi
  python -m pytest tests/test_gpu_lloyd.py -m gpu -x -q -k "whole_gcvt or baseline_sizes or full" 2>&1 | tail -1
  python bench.py --no-cpu --steps 300 --warmup 20 --e2e-iters 50 > $OUT/ab_$v.json 2> $OUT/ab_$v.err
  python - <<PY
import json
d=json.loads(open("$OUT/ab_$v.json").read().strip().splitlines()[-1])
print("$v", round(d["value"],1), "it/s  e2e", round(d["e2e"]["value"],1), {k:round(x*1e3,1) for k,x in d["roofline"]["stages_ms_per_step"].items()})
PY
done
