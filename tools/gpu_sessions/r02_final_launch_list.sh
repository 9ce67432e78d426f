#!/bin/bash
# launch list of the default bench on the final code (durations only, no replay)
mkdir -p gpurun_out
timeout 88 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_r02f.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-plummer --no-parity > gpurun_out/launches_r02f_bench.log 2>&1; echo "launch list rc=$?"
wc -l gpurun_out/launches_r02f.csv
