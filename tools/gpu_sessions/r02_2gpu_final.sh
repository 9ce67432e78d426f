#!/bin/bash
# GPU call 8 (2 GPUs): full suite on the final code, bench on 1 and 2 GPUs
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_final_2gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_final_2gpu.log; tail -4 gpurun_out/r2_pytest_final_2gpu.log
timeout 900 python bench.py > gpurun_out/r2_bench_d_1gpu.json 2> gpurun_out/r2_bench_d_1gpu.err; echo "bench1 rc=$?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29601 bench.py --gpus 2 > gpurun_out/r2_bench_d_2gpu.json 2> gpurun_out/r2_bench_d_2gpu.err; echo "bench2 rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/r2_bench_d_1gpu.json", "gpurun_out/r2_bench_d_2gpu.json"):
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(f, "naive %.4g frac %.4f" % (d["value"], d["roofline"]["frac"]))
    for k in ("bh", "bh_plummer", "bh_plummer_static_slices", "bh_config3"):
        b = d.get(k)
        if b: print("  ", k, b.get("value"), b.get("ms_per_step"), (b.get("phases_ms") or {}).get("Octree creation"), (b.get("parity") or {}).get("ok"), b.get("checksum"), b.get("walk_max_over_mean"), b.get("error"))
    print("  ", d.get("config1"))
PY
