#!/bin/bash
# 8-GPU session of round 2: multi-GPU parity (worlds 4 and 8), bench on 8 GPUs, BASELINE config 5 (N = 2^26, theta 0.2, energy)
mkdir -p gpurun_out
{ nvidia-smi -L; free -g; nproc; } > gpurun_out/r2_8gpu_box.txt 2>&1
timeout 1200 python -m pytest tests/test_gpu_multi.py -q -k "4-True or 8-True or 8-False" > gpurun_out/r2_pytest_8gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_8gpu.log
tail -6 gpurun_out/r2_pytest_8gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2_bench_8gpu.json 2> gpurun_out/r2_bench_8gpu.err; echo "bench8 rc=$?"
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2_bench_8gpu.json").read().strip().splitlines()[-1])
    print("naive %.4g int/s" % d["value"], "e2e", d["e2e"]["value"])
    for k in ("bh", "bh_plummer", "bh_plummer_static_slices"):
        b = d.get(k) or {}
        print(k, "steps/s", b.get("value"), "ms", b.get("ms_per_step"), "walk/rank", b.get("walk_ms_per_rank"), "max/mean", b.get("walk_max_over_mean"),
              "phases", b.get("phases_ms"), "checksum", b.get("checksum"), "parity", (b.get("parity") or {}).get("ok"), "p2p", b.get("p2p"), b.get("error"))
except Exception as e:
    print("bench8 unreadable", e)
PY
tail -3 gpurun_out/r2_bench_8gpu.err
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29621 tools/config5.py --out gpurun_out/r2_config5_8gpu.json > gpurun_out/r2_config5_8gpu.log 2>&1; echo "config5 rc=$?"
tail -4 gpurun_out/r2_config5_8gpu.log | cut -c1-3000
