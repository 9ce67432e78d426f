#!/bin/bash
# final 8-GPU bench of round 2 (default flags, as the driver launches it), plus the small-N timing on one GPU
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 8 > gpurun_out/r2_bench_d_8gpu.json 2> gpurun_out/r2_bench_d_8gpu.err; echo "bench8 rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2_bench_d_8gpu.json").read().strip().splitlines()[-1])
print("naive %.4g int/s e2e %.4g" % (d["value"], d["e2e"]["value"]))
for k in ("bh", "bh_plummer", "bh_plummer_static_slices"):
    b = d.get(k) or {}
    print(k, "steps/s", b.get("value"), "ms", b.get("ms_per_step"), "walk/rank", b.get("walk_ms_per_rank"), "max/mean", b.get("walk_max_over_mean"),
          "tree", (b.get("phases_ms") or {}).get("Octree creation"), "checksum", b.get("checksum"), "parity", (b.get("parity") or {}).get("ok"), "p2p", b.get("p2p"), b.get("error"))
PY
timeout 300 python tools/dev_small_n.py 4000 2>&1 | tee gpurun_out/r2_small_n.log
