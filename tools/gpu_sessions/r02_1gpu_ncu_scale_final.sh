#!/bin/bash
# GPU call 6 (1 GPU): ncu evidence of the round-2 code, 2^26 sampled parity, full suite, final 1-GPU bench
mkdir -p gpurun_out
# (1) launch list of the default bench (no replay: durations only)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_r02.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-plummer --no-parity > gpurun_out/launches_r02_bench.log 2>&1; echo "launch list rc=$?"
# (2) full captures of the dominant kernels
timeout 600 ncu --set full --clock-control none --import-source on -k regex:naive_accel -s 1 -c 1 -o gpurun_out/prof_naive_r02 -f \
    python bench.py --steps 1 --warmup 3 --no-bh --no-cpu > gpurun_out/prof_naive_r02.log 2>&1; echo "ncu naive rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"bh_traverse|os_scatter|com_levels|reorder|emit_kernel|keys_kernel" -s 40 -c 24 -o gpurun_out/prof_bh_r02 -f \
    python tools/dev_ab_step.py 16777216 uniform_sphere 0.5 4 > gpurun_out/prof_bh_r02.log 2>&1; echo "ncu bh rc=$?"
# (3) sampled parity at N = 2^26, theta = 0.2 (config 5's size)
NB_SCALE_TESTS=1 timeout 1500 python -m pytest tests/test_gpu_scale.py -q -s -k 64m > gpurun_out/r2_scale64m.log 2>&1; tail -4 gpurun_out/r2_scale64m.log
# (4) whole suite and the final 1-GPU bench
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_final_1gpu.log 2>&1; tail -3 gpurun_out/r2_pytest_final_1gpu.log
timeout 900 python bench.py > gpurun_out/r2_bench_final_1gpu.json 2> gpurun_out/r2_bench_final_1gpu.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference > gpurun_out/r2_bench_final_reference.json 2> gpurun_out/r2_bench_final_reference.err; echo "bench ref rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2_bench_final_1gpu.json").read().strip().splitlines()[-1])
print("naive %.4g frac %.4f micro %.4f e2e %.4g" % (d["value"], d["roofline"]["frac"], d["roofline"]["frac_of_microbenchmark"], d["e2e"]["value"]))
for k in ("bh", "bh_plummer", "bh_config3"):
    b = d.get(k) or {}
    print(k, b.get("value"), b.get("ms_per_step"), (b.get("roofline") or {}).get("frac"), (b.get("parity") or {}).get("ok"), (b.get("e2e") or {}).get("value"), b.get("error"))
print(d.get("config1"))
PY
ls -la gpurun_out/*.ncu-rep
