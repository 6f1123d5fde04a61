#!/bin/bash
# Multi-GPU measurement on one box (run under `gpurun --gpus 8`): correctness of the sharded loop across real
# processes, the headline config at N ranks, and BASELINE configs[3] (32768^2 / 1M sites) at 8 ranks.
# Every launch is bounded by `timeout`; outputs go to gpurun_out/.
set -u
OUT=gpurun_out
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
nvidia-smi -L | head -8 > $OUT/scale_gpus.txt
free -g | head -2 >> $OUT/scale_gpus.txt
NG=$(nvidia-smi -L | wc -l)
echo "gpus=$NG" >> $OUT/scale_gpus.txt
timeout 300 $TR --nproc-per-node $NG --master-port 29511 tools/verify_multi.py 4096 20000 25 p2p > $OUT/vm_n${NG}_p2p.out 2> $OUT/vm_n${NG}_p2p.err
echo "verify rc=$?"; tail -1 $OUT/vm_n${NG}_p2p.out
for N in ${SCALE_NS:-8 4}; do
  [ $N -le $NG ] || continue
  timeout 400 $TR --nproc-per-node $N --master-port $((29520+N)) bench.py --gpus $N --steps 300 --warmup 20 --no-cpu > $OUT/bench_s2_n${N}.json 2> $OUT/bench_s2_n${N}.err
  echo "bench N=$N rc=$?"; tail -c 600 $OUT/bench_s2_n${N}.json | head -c 300; echo
done
MEM=$(free -g | awk '/Mem:/{print $7}')
if [ "${SKIP_C4:-0}" = "0" ] && [ "$MEM" -gt 250 ]; then
  timeout 600 $TR --nproc-per-node $NG --master-port 29540 bench.py --gpus $NG --grid 32768 --sites 1000000 --steps 60 --warmup 6 --e2e-iters 10 --no-cpu > $OUT/bench_s2_c4_n${NG}.json 2> $OUT/bench_s2_c4_n${NG}.err
  echo "C4 N=$NG rc=$?"; head -c 400 $OUT/bench_s2_c4_n${NG}.json; echo
else
  echo "C4 skipped (free host memory ${MEM} GB)"
fi
