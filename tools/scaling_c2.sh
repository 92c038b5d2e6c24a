#!/bin/bash
# C2 weak scaling on one multi-GPU box: N = 1, 2, 4, 8 (or the subset given as arguments).
mkdir -p gpurun_out
for N in ${@:-1 2 4 8}; do
  port=$((29700 + RANDOM % 200))
  if [ "$N" = "1" ]; then
    python bench.py --gpus 1 --steps 30 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/scale_v6_N$N.json
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port \
      bench.py --gpus $N --steps 30 --warmup 3 --no-cpu-baseline 2>&1 | grep '^{"metric"' | tail -1 > gpurun_out/scale_v6_N$N.json
  fi
  python -c "
import json
d = json.load(open('gpurun_out/scale_v6_N$N.json'))
print('N=%d: %.1f instance-it/s, %.3f ms/step, e2e %.1f, clocks %s' % (d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks']))"
done
