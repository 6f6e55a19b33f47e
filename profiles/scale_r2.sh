#!/bin/bash
# The driver's own multi-GPU command at the N given (run with: gpurun --gpus N -- bash profiles/scale_r2.sh N):
# the full bench -- headline step with the peer-mailbox exchange and its in-run parity assertion, sharded
# post-process + detection all-gather, D7 -- one JSON line into gpurun_out/scale_N.json.
N=${1:-2}
if [ "$N" = 1 ]; then
  python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/scale_1.json 2> gpurun_out/scale_1.err
else
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$N \
      bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/scale_$N.json 2> gpurun_out/scale_$N.err || echo "N=$N failed"
fi
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/scale_$N.json").read().strip().splitlines()[-1])
    print("N", d["n_gpus"], "ms/step", d["ms_per_step"], "images/s", d["value"], d["config"].get("launch"))
    print("exchange:", d["config"].get("exchange"))
    print("exchange_check:", d["config"].get("exchange_check"))
    for k, v in d.get("workloads", {}).items():
        print(k, v.get("ms_per_step"), v.get("value"), v.get("launch"), v.get("gather"))
except Exception as e:
    print("ERR", e)
    print(open("gpurun_out/scale_$N.err").read()[-3000:])
PY
