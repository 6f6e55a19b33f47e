#!/bin/bash
# weak-scaling check of the headline step at N = 4 and 8 (run with: gpurun --gpus 8 -- bash profiles/scale8.sh)
for n in 4 8; do
  timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n \
      bench.py --gpus $n --steps 50 --warmup 5 --no-extra --no-cpu-baseline > gpurun_out/s$n.json 2>gpurun_out/s$n.err || echo "N=$n failed"
done
timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 \
    bench.py --impl reference --gpus 8 --steps 2 --warmup 1 > gpurun_out/s8ref.json 2>gpurun_out/s8ref.err || echo "ref failed"
python - <<PY
import json
for f in ("gpurun_out/s4.json", "gpurun_out/s8.json", "gpurun_out/s8ref.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d.get("n_gpus"), d.get("ms_per_step"), d.get("value"), d["config"].get("launch"), d.get("impl"))
    except Exception as e:
        print(f, "ERR", e)
PY
