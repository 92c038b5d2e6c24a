#!/bin/bash
# Weak-scaling sweep on one 8-GPU box: C2 at N = 1, 2, 4, 8 and the C5 Hilbert-dimension sweep at N = 8.
# Usage (gpurun --gpus 8): bash tools/scaling_sweep.sh
mkdir -p gpurun_out
run() {  # N workload steps
  local N=$1 W=$2 S=$3 port=$((29700 + RANDOM % 200))
  if [ "$N" = "1" ]; then
    python bench.py --gpus 1 --workload $W --steps $S --warmup 2 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/scale_${W}_N${N}.json
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port \
      bench.py --gpus $N --workload $W --steps $S --warmup 2 --no-cpu-baseline 2>&1 | grep '^{"metric"' | tail -1 > gpurun_out/scale_${W}_N${N}.json
  fi
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/scale_${W}_N${N}.json"))
    print("${W} N=${N}: %.1f instance-it/s, %.2f ms/step, clocks %s" % (d["value"], d["ms_per_step"], d["clocks"]))
except Exception as e:
    print("${W} N=${N}: FAILED", e)
PY
}
for N in 1 2 4 8; do run $N C2 30; done
run 8 C5n8 5
run 8 C5n16 5
run 8 C5n32 5
run 8 C5n64 3
