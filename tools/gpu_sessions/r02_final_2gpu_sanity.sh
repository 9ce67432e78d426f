#!/bin/bash
# Last GPU call of round 2 (2 GPUs, ~2 minutes of budget): the sharded ranks with peer stores must still equal one GPU
# bit for bit with the 48-warp walk, then a short 2-GPU bench line
mkdir -p gpurun_out
timeout 70 python -m pytest tests/test_gpu_multi.py -q -k "test_ranks_match_single_gpu and 2-True" > gpurun_out/r2f_pytest_2gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2f_pytest_2gpu.log; tail -n 5 gpurun_out/r2f_pytest_2gpu.log
timeout 45 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --no-cpu --no-plummer --no-parity > gpurun_out/r2f_bench_2gpu.json 2> gpurun_out/r2f_bench_2gpu.err; echo "bench2 rc=$?"
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2f_bench_2gpu.json").read().strip().splitlines()[-1])
    b = d.get("bh") or {}
    print("naive %.4g" % d["value"], "bh", b.get("value"), b.get("ms_per_step"), b.get("phases_ms"), b.get("checksum"), b.get("walk_ms_per_rank"))
except Exception as e:
    print("no bench line:", e)
PY
