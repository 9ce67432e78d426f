#!/bin/bash
# GPU call 1 of round 2: parity suite on the new build, then A/B of the Barnes-Hut step against the round-1 library
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest1.log
tail -5 gpurun_out/r2_pytest1.log
for n in 16777216 1048576; do
  gen=uniform_sphere; [ $n = 1048576 ] && gen=plummer
  NB_LIB=$PWD/n-body-simulation_b200/libnbody_b200_r1.so timeout 300 python tools/dev_ab_step.py $n $gen 0.5 6 >> gpurun_out/r2_ab1.log 2>&1
  timeout 300 python tools/dev_ab_step.py $n $gen 0.5 6 >> gpurun_out/r2_ab1.log 2>&1
done
cat gpurun_out/r2_ab1.log
