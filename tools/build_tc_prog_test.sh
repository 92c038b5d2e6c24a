#!/bin/bash
# builds tools/tc_prog_test.bin (the hardware check of the tcgen05/TMA program engine) for sm_100a
set -e
cd "$(dirname "$0")/.."
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -I include -I quantum-optimal-control_b200/csrc \
  tools/tc_prog_test.cu quantum-optimal-control_b200/csrc/qoc_tc_f16.cu quantum-optimal-control_b200/csrc/qoc_tc_small.cu quantum-optimal-control_b200/csrc/qoc_tc_pair.cu -o tools/tc_prog_test.bin
