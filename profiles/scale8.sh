#!/bin/bash
# weak-scaling check of the headline step at N = 8 (run with: gpurun --gpus 8 -- bash profiles/scale8.sh)
for n in ${NS:-8}; do
  timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n \
      bench.py --gpus $n --steps 50 --warmup 5 --no-extra --no-cpu-baseline > gpurun_out/s$n.json 2>gpurun_out/s$n.err || echo "N=$n failed"
done
python - <<PY
import json, os
for n in os.environ.get("NS", "8").split():
    f = "gpurun_out/s%s.json" % n
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d.get("n_gpus"), d.get("ms_per_step"), d.get("value"), d["config"].get("launch"), d["config"].get("exchange"), d["config"].get("loss_out"))
    except Exception as e:
        print(f, "ERR", e)
PY
