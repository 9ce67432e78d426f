#!/bin/bash
# GPU call (1 GPU): A/B of the 36-instruction cursor step (exact per-depth thresholds, 256-bit node load, register-resident
# bases) against the previous library, with 40 and 48 resident warps; phase timers of the build (4-pass sort); full suite
mkdir -p gpurun_out
L=gpurun_out/r2_walk36_ab.log
: > $L
for lib in build_ab/base/libnbody_b200.so "" build_ab/occ6/libnbody_b200.so; do
  NB_LIB=${lib:+$PWD/$lib} timeout 200 python tools/dev_walk_sweep.py 16777216 0 128 >> $L 2>&1
done
for lib in build_ab/base/libnbody_b200.so ""; do
  NB_LIB=${lib:+$PWD/$lib} timeout 200 python tools/dev_ab_step.py 16777216 uniform_sphere 0.5 6 >> $L 2>&1
done
cat $L
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_walk36.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_walk36.log; tail -n 15 gpurun_out/r2_pytest_walk36.log
