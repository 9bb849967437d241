#!/bin/bash
# round-2 GPU call T: ncu --set full of the shipped frame kernels (-> profiles/traffic.json), set-up kernel list, final bench lines
mkdir -p gpurun_out
ncu --set full --import-source on --clock-control none -k regex:"k_ptrace|k_shade|k_trace|k_accum|k_ascent" -s 14 -c 10 -o gpurun_out/r02t_full -f python tools/ab_bench.py --frames 8 --warmup 8 --repeat 1 > gpurun_out/r02t_full.log 2>&1
tail -1 gpurun_out/r02t_full.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r02t_setup_launches.csv python tools/ab_bench.py --frames 4 --warmup 4 --repeat 1 > /dev/null 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open("gpurun_out/r02t_setup_launches.csv")))
h=[i for i,r in enumerate(rows) if r and r[0]=="ID"][0]
c=rows[h]; ki=c.index("Kernel Name"); vi=c.index("Metric Value")
for r in rows[h+1:h+30]:
    if len(r)>vi: print("SETUP", r[0], r[ki][:50], r[vi])
PY
python bench.py --steps 20 --warmup 5 > gpurun_out/r02t_bench20.json 2> gpurun_out/r02t_bench.err; tail -c 300 gpurun_out/r02t_bench20.json; echo
python - <<'PY'
import json
for line in open("gpurun_out/r02t_bench20.json"):
    if line.startswith("{"):
        d=json.loads(line); print("bench20", d["value"], d["ms_per_step"], d["roofline"]["frac"], d["e2e"])
PY
