#!/bin/bash
# Scaling lines of one workload on one box: builds the tree once (snapshot in /tmp), then runs bench.py for every
# "N:shard" given.   tools/run_scaling.sh cfg4 r02_bench_cfg4 "1:query 2:query 2:store"
wl=$1; tag=$2; runs=$3
mkdir -p gpurun_out
snap=/tmp/${wl}.cwb
port=29500
for r in $runs; do
  n=${r%%:*}; shard=${r##*:}
  out=gpurun_out/${tag}_n${n}_${shard}.json
  if [ "$n" = "1" ]; then
    timeout 1500 python bench.py --workload $wl --snapshot $snap --no-cpu-baseline --steps 3 --warmup 3 > $out 2> gpurun_out/${tag}_n${n}_${shard}.err
  else
    port=$((port+1))
    timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port \
      bench.py --gpus $n --workload $wl --snapshot $snap --shard $shard --no-cpu-baseline --steps 3 --warmup 3 > $out 2> gpurun_out/${tag}_n${n}_${shard}.err
  fi
  echo "== $r exit $? $(python -c "
import json,sys
try:
    j=json.load(open('$out')); print(round(j['value']), 'q/s', round(j['ms_per_step'],2), 'ms', 'e2e', round(j['e2e']['value']), 'recall', j['recall_at_k'], 'flagged', (j.get('fused_stats') or {}).get('flagged'))
except Exception as e: print('no line', e)")"
  tail -2 gpurun_out/${tag}_n${n}_${shard}.err | cut -c1-300
done
