#!/bin/bash
# All BASELINE.json configurations through bench.py on one GPU (3-5 timed steps each); one JSON line per workload.
mkdir -p gpurun_out
run() { # name, extra args
  timeout 600 python bench.py --workload $1 $2 --no-cpu-baseline > gpurun_out/sec_$1${3}.json 2> gpurun_out/sec_$1${3}.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/sec_$1${3}.json")); r = d["roofline"]; k = r["kernel_ms_per_step"]
    print("| $1${3} | %s | %.1f | %.2f | %.2f | %.2f | %.2f | %.2f | %.3f | %.3f | %.1f |" % (d["config"]["workload"], d["value"], d["ms_per_step"], k["expm"], k["chain"], k["costate"], k["grad"], r["frac"] or 0, r["executed_frac_of_peak"] or 0, r["whole_step_alg_tflops"]))
except Exception as e:
    print("$1${3} FAILED", e)
PY
}
run C2 "--steps 30 --warmup 3"
run C2 "--steps 30 --warmup 3 --dtype tf32x3" _tf32x3
run C3 "--steps 3 --warmup 1"
run C5n8 "--steps 5 --warmup 2"
run C5n16 "--steps 5 --warmup 2"
run C5n32 "--steps 5 --warmup 2"
run C5n64 "--steps 3 --warmup 1"
run C4 "--steps 2 --warmup 1 --batch 16" _B16
