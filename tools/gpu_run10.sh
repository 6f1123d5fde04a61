#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
( time timeout 900 python -m pytest tests -m gpu -x -q --durations=3 ) > $OUT/pytest_gpu.log 2>&1
tail -6 $OUT/pytest_gpu.log
SRM_TRACE=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --c4-steps 0 > $OUT/bench10.json 2> $OUT/bench10.err; grep srm_gcvt $OUT/bench10.err | tail -6; python -c "
import json;d=json.loads([l for l in open('$OUT/bench10.json') if l.startswith('{')][-1]);print('value',d['value'],'e2e',d['e2e']['value'],d['e2e']['ms_per_call'])"
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_write.sum --clock-control none -k regex:k_expand\|k_prefix -c 4 --csv --log-file $OUT/launches10.csv python -c "
import sys; sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import bench, surface_remesher_b200 as S
d,m,v=bench.make_inputs(8192,100000,False)
b=v.copy(); S.gCVT(b,d,m,8192,1,5); b=v.copy(); S.gCVT(b,d,m,8192,1,5)" > /dev/null 2>&1; grep -E "k_expand|k_prefix" $OUT/launches10.csv | cut -d, -f5,13-15 | head -8
