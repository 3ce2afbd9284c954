#!/bin/bash
# Diagnostic build with k_detect phase clocks (-DSPVO_PHASE_TIMING) -> scripts/_diag/libspvo_timing.so (see scripts/detect_phases.py)
set -e
cd "$(dirname "$0")/../superpoint-stereo-visual-odometry_b200/csrc"
mkdir -p /tmp/spvo_diag ../../scripts/_diag
for f in api decode match preprocess match_tc; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -Xcompiler -fPIC -I../../include -DSPVO_PHASE_TIMING -w -c -o /tmp/spvo_diag/$f.o $f.cu &
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../scripts/_diag/libspvo_timing.so /tmp/spvo_diag/*.o
