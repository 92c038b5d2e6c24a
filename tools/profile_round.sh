#!/bin/bash
# Round profile pass (run under gpurun, one GPU): launch list, ncu --set full of the hot kernels, final bench line.
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_v7.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/launches_v7.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_expm_mma|k_segprod|k_vec_sweep|k_chain_mma' -s 5 -c 5 \
    -o gpurun_out/prof_v7 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/prof_v7.log 2>&1
python bench.py --steps 100 --warmup 5 > gpurun_out/bench_v7.json 2> gpurun_out/bench_v7.err
cat gpurun_out/bench_v7.json
