#!/bin/bash
# ramped first batch A/B on the driver's line
mkdir -p gpurun_out
for rep in 1 2; do for r in 0 1; do
  F3D_B200_RAMP=$r python bench.py --gpus 1 --steps 20 --warmup 5 --no-secondary --no-cpu-baseline --no-rows > gpurun_out/ramp_$r.json 2> gpurun_out/ramp.err
  python - <<PY
import json
d=[json.loads(l) for l in open("gpurun_out/ramp_$r.json") if l.startswith("{")][-1]
print("ramp $r steps 20: ms/step", round(d["ms_per_step"],4), "value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "call_ms", round(d["e2e"]["call_ms"],2))
PY
done; done
for r in 0 1; do F3D_B200_RAMP=$r python tools/ab_bench.py 2>&1 | tail -1; done
