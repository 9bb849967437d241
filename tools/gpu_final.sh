#!/bin/bash
# dress rehearsal of the driver's round-end sequence on one B200: smoke(), the -m gpu suite, both bench arms
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/final_ref.json 2> gpurun_out/final.err
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/final_bench.json 2>> gpurun_out/final.err
python - <<'PY'
import json
r=[json.loads(l) for l in open("gpurun_out/final_ref.json") if l.startswith("{")][-1]
d=[json.loads(l) for l in open("gpurun_out/final_bench.json") if l.startswith("{")][-1]
print("reference", r["value"], "| ours value", d["value"], "ms", d["ms_per_step"], "frac", d["roofline"]["frac"], "traffic", d["roofline"]["traffic"], "| e2e", d["e2e"]["value"], "ratio e2e", d["e2e"]["value"]/r["value"])
print("rows", {k:(v.get("roofline") or v.get("roofline_issue") or v.get("error")) for k,v in d["widened_rows"].items()})
print("fast", d.get("fast_numerics"))
PY
