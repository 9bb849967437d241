#!/bin/bash
# round-2 GPU call J: split primary pass A/B
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_multigpu.py -m gpu -x -q 2>&1 | tail -3
python tools/ab_bench.py 2>&1 | tail -1
F3D_B200_NO_SPLIT=1 python tools/ab_bench.py 2>&1 | tail -1
F3D_B200_BATCH=8 python tools/ab_bench.py 2>&1 | tail -1
echo "--- 1/8"
python tools/ab_bench.py --part 0/8 2>&1 | tail -1
F3D_B200_NO_SPLIT=1 python tools/ab_bench.py --part 0/8 2>&1 | tail -1
F3D_B200_BATCH=8 python tools/ab_bench.py --part 0/8 2>&1 | tail -1
F3D_B200_BATCH=8 F3D_B200_TRACE_CTAS=4 python tools/ab_bench.py --part 0/8 2>&1 | tail -1
F3D_B200_BATCH=8 F3D_B200_TRACE_CTAS=2 python tools/ab_bench.py --part 0/8 2>&1 | tail -1
echo "--- 1/4, 1/2"
python tools/ab_bench.py --part 3/4 2>&1 | tail -1
python tools/ab_bench.py --part 1/2 2>&1 | tail -1
ncu --metrics gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"k_ptrace|k_shade|k_trace|k_accum|k_ascent" -s 40 -c 16 --csv --log-file gpurun_out/r02j_launches.csv python tools/ab_bench.py --frames 8 --warmup 8 --repeat 1 > /dev/null 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open("gpurun_out/r02j_launches.csv")))
h=[i for i,r in enumerate(rows) if r and r[0]=="ID"][0]
c=rows[h]; ki=c.index("Kernel Name"); mi=c.index("Metric Name"); vi=c.index("Metric Value")
agg={}
for r in rows[h+1:]:
    if len(r)>vi: agg.setdefault(r[ki][:40],{}).setdefault(r[mi],[]).append(float(r[vi].replace(",","")))
for k,v in agg.items():
    print("NCU", k, len(v["gpu__time_duration.sum"]), {m: round(sum(x)/len(x),2) for m,x in v.items()})
PY
