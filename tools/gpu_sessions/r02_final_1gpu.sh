#!/bin/bash
# Final 1-GPU call of round 2: whole GPU suite on the final library, default bench, ncu capture of the walk
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/r2f_pytest_1gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2f_pytest_1gpu.log; tail -n 4 gpurun_out/r2f_pytest_1gpu.log
timeout 600 python bench.py > gpurun_out/r2f_bench_1gpu.json 2> gpurun_out/r2f_bench_1gpu.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2f_bench_1gpu.json").read().strip().splitlines()[-1])
print("naive %.4g frac %.4f e2e %.4g" % (d["value"], d["roofline"]["frac"], d["e2e"]["value"]))
for k in ("bh", "bh_plummer", "bh_config3"):
    b = d.get(k) or {}
    print(k, b.get("value"), b.get("ms_per_step"), (b.get("roofline") or {}).get("frac"), (b.get("parity") or {}).get("ok"), (b.get("e2e") or {}).get("value"), b.get("phases_ms"), b.get("error"))
print(d.get("config1"))
PY
timeout 300 ncu --set full --clock-control none --import-source on -k regex:bh_traverse -s 3 -c 1 -o gpurun_out/prof_walk_r02f -f \
    python tools/dev_ab_step.py 16777216 uniform_sphere 0.5 4 > gpurun_out/prof_walk_r02f.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/*.ncu-rep | tail -n 2
