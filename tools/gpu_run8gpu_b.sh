#!/bin/bash
# 8-GPU session 2: parity tests on 2/4/8 ranks, benches at N = 8 and 4 (8192^2 + C4), C4 with work-balanced bands
OUT=gpurun_out; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
( time timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q ) > $OUT/pytest_multi8.log 2>&1; tail -3 $OUT/pytest_multi8.log
timeout 900 $TR --nproc-per-node 8 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 5 > $OUT/bench_n8.json 2> $OUT/bench_n8.err; tail -2 $OUT/bench_n8.err
timeout 600 $TR --nproc-per-node 8 --master-port 29522 bench.py --gpus 8 --only-c4 --bands balanced > $OUT/c4_n8_balanced.json 2> $OUT/c4_n8_balanced.err
timeout 600 $TR --nproc-per-node 4 --master-port 29523 bench.py --gpus 4 --steps 20 --warmup 5 > $OUT/bench_n4.json 2> $OUT/bench_n4.err
python - <<'PY'
import json
def last(f):
    try: return json.loads([l for l in open(f"gpurun_out/{f}") if l.startswith('{')][-1])
    except Exception as e: return {"ERR": str(e)}
for f in ("bench_n8.json","bench_n4.json"):
    d=last(f)
    if "ERR" in d: print(f,d); continue
    print(f,'value',round(d['value']),'ms/step',round(d['ms_per_step'],4),'e2e',round(d['e2e']['value']),round(d['e2e']['ms_per_call'],1),'launches',d['gpu_launches'],'parity',d['parity'])
    print('   stages',d['config'].get('stages_ms_per_step_rank0')); print('   c4',d.get('c4'))
print('c4 balanced', last("c4_n8_balanced.json"))
PY
