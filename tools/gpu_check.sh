#!/bin/bash
# developer helper, usage (under gpurun, from the repo root): bash tools/gpu_check.sh <tag> -- parity tests, bench line, per-kernel ncu durations
TAG=${1:-x}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 64 --warmup 8 --no-cpu-baseline --no-e2e > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_$TAG.json"))
print("BENCH", "$TAG", "ms/frame %.4f" % d["ms_per_step"], "Mrays/s %.1f" % d["value"], "frac %.4f" % d["roofline"]["frac"], "nodes/ray %.2f" % d["config"]["nodes_per_ray"])
PY
tail -2 gpurun_out/bench_$TAG.err
ncu --metrics gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum --clock-control none -k regex:"k_primary|k_trace|k_accum" -s 40 -c 8 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 4 --warmup 8 --no-cpu-baseline --no-e2e > /dev/null 2>&1
python - <<PY
import csv
rows=list(csv.reader(open("gpurun_out/launches_$TAG.csv")))
h=[i for i,r in enumerate(rows) if r and r[0]=="ID"][0]
c=rows[h]; ki=c.index("Kernel Name"); mi=c.index("Metric Name"); vi=c.index("Metric Value")
agg={}
for r in rows[h+1:]:
    if len(r)>vi: agg.setdefault(r[ki][:40],{}).setdefault(r[mi],[]).append(float(r[vi].replace(",","")))
for k,v in agg.items():
    print("NCU", k, {m: round(sum(x)/len(x),2) for m,x in v.items()})
PY
