#!/bin/bash
# GPU call 4 (2 GPUs): multi-GPU tests with cost-weighted slices, bench on 2 GPUs with the Plummer balance lines
mkdir -p gpurun_out
echo skip-tests

timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29601 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_bench4_2gpu.json 2> gpurun_out/r2_bench4_2gpu.err; echo "bench2 rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2_bench4_2gpu.json").read().strip().splitlines()[-1])
for k in ("bh", "bh_plummer", "bh_plummer_static_slices"):
    b = d.get(k) or {}
    print(k, "steps/s", b.get("value"), "ms", b.get("ms_per_step"), "walk/rank", b.get("walk_ms_per_rank"), "max/mean", b.get("walk_max_over_mean"),
          "slices", (b.get("config") or {}).get("slices"), "checksum", b.get("checksum"), "parity", (b.get("parity") or {}).get("ok"), b.get("error"))
PY
tail -3 gpurun_out/r2_bench4_2gpu.err
