#!/bin/bash
# GPU call 3 (1 GPU): naive kernel sweep (no-producer-warp forms), walk with CTA-synchronous rounds
mkdir -p gpurun_out
timeout 600 python tools/dev_naive_sweep.py 524288 0,12,13,14,15,16,17,18,19 256 > gpurun_out/r2_naive_sweep.log 2>&1
timeout 300 python tools/dev_naive_sweep.py 524288 0,12,19 512,128 >> gpurun_out/r2_naive_sweep.log 2>&1
cat gpurun_out/r2_naive_sweep.log
timeout 600 python tools/dev_walk_sweep.py 16777216 50,51 128,256 > gpurun_out/r2_walk_sync.log 2>&1
cat gpurun_out/r2_walk_sync.log
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "tree or sort or dense or deep or advance or massless or coincident or pool" > gpurun_out/r2_pytest3.log 2>&1; tail -3 gpurun_out/r2_pytest3.log
timeout 300 python tools/dev_ab_step.py 16777216 uniform_sphere 0.5 6 > gpurun_out/r2_ab3.log 2>&1; cat gpurun_out/r2_ab3.log
