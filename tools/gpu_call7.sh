#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/r2_build_ab.log
for it in 8 12 16; do NB_OS_ITEMS=$it timeout 200 python tools/dev_build_ab.py 16777216 0 2>&1 | tail -1 >> gpurun_out/r2_build_ab.log; done
NB_EMIT_PER_BODY=1 timeout 200 python tools/dev_build_ab.py 16777216 0 2>&1 | tail -1 >> gpurun_out/r2_build_ab.log
timeout 300 python tools/dev_build_ab.py 16777216 0,2,3,4 2>&1 | tail -4 >> gpurun_out/r2_build_ab.log
cat gpurun_out/r2_build_ab.log
NB_OS_ITEMS=12 timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "tree or sort or dense or deep" 2>&1 | tail -3
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_reference.py -x -q 2>&1 | tail -3
