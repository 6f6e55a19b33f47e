#!/bin/bash
# N=1 vs N=2 weak-scaling check of the headline step (run with: gpurun --gpus 2 -- bash profiles/scale2.sh)
python bench.py --steps 50 --warmup 5 --no-extra --no-cpu-baseline > gpurun_out/s1.json 2>gpurun_out/s1.err
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 50 --warmup 5 --no-extra --no-cpu-baseline > gpurun_out/s2.json 2>gpurun_out/s2.err || echo "N=2 failed"
python - <<PY
import json
for f in ("gpurun_out/s1.json", "gpurun_out/s2.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["n_gpus"], d["ms_per_step"], d["value"], d["config"].get("launch"), d["config"].get("loss_out"))
    except Exception as e:
        print(f, "ERR", e)
PY
