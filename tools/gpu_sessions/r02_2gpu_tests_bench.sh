#!/bin/bash
# GPU call 2 of round 2 (2 GPUs): whole parity suite incl. multi-GPU (IPC peer stores and NCCL fallback) and the 2^24
# scale test, then the bench on 1 and 2 GPUs
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2_call2_gpus.txt
timeout 1500 python -m pytest tests -m gpu -q --durations=15 > gpurun_out/r2_pytest2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest2.log
tail -30 gpurun_out/r2_pytest2.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_1gpu.json 2> gpurun_out/r2_bench_1gpu.err; echo "bench1 rc=$?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29601 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_bench_2gpu.json 2> gpurun_out/r2_bench_2gpu.err; echo "bench2 rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/r2_bench_1gpu.json", "gpurun_out/r2_bench_2gpu.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        bh = d.get("bh") or {}
        print(f, "naive %.4g int/s frac %.3f (micro %.3f) | bh %.3f steps/s %.2f ms walk %.2f frac %.3f parity %s p2p %s e2e %s" % (
            d["value"], d["roofline"]["frac"], d["roofline"]["frac_of_microbenchmark"], bh.get("value", 0), bh.get("ms_per_step", 0),
            (bh.get("phases_ms") or {}).get("Acceleration Kernel Time", 0), (bh.get("roofline") or {}).get("frac", 0),
            (bh.get("parity") or {}).get("ok"), bh.get("p2p"), (bh.get("e2e") or {}).get("value")))
        print("   phases", bh.get("phases_ms"), "checksum", bh.get("checksum"), "err", bh.get("error"))
        if "bh_config3" in d:
            c = d["bh_config3"]; print("   config3", c.get("value"), c.get("ms_per_step"), c.get("cpu_baseline"), c.get("error"))
        if "config1" in d:
            print("   config1", d["config1"])
    except Exception as e:
        print(f, "unreadable", e)
PY
tail -5 gpurun_out/r2_bench_1gpu.err gpurun_out/r2_bench_2gpu.err
