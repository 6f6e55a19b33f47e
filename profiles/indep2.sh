#!/bin/bash
# two INDEPENDENT single-GPU runs of the headline step, first one after the other, then concurrently:
# separates "the second GPU / two busy GPUs are slower" from "the exchange costs time"
A="--steps 200 --warmup 5 --no-extra --no-cpu-baseline"
CUDA_VISIBLE_DEVICES=1 python bench.py $A > gpurun_out/i_gpu1_alone.json 2>/dev/null
CUDA_VISIBLE_DEVICES=0 python bench.py $A > gpurun_out/i_gpu0_both.json 2>/dev/null &
P0=$!
CUDA_VISIBLE_DEVICES=1 python bench.py $A > gpurun_out/i_gpu1_both.json 2>/dev/null &
P1=$!
wait $P0 $P1
python - <<PY
import json
for f in ("i_gpu1_alone", "i_gpu0_both", "i_gpu1_both"):
    d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
    print(f, d["ms_per_step"], d["roofline"]["kernel_ms"], d["clocks"])
PY
