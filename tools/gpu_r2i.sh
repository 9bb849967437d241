#!/bin/bash
# round-2 GPU call I: small-partition experiments (stand-alone 1/8 image)
mkdir -p gpurun_out
for c in 2 3 4 5 6; do echo "TRACE_CTAS=$c"; F3D_B200_TRACE_CTAS=$c python tools/ab_bench.py --part 0/8 2>&1 | tail -1; done
echo "SETS=1"; F3D_B200_SETS=1 python tools/ab_bench.py --part 0/8 2>&1 | tail -1
echo "full frame TRACE_CTAS=4"; F3D_B200_TRACE_CTAS=4 python tools/ab_bench.py 2>&1 | tail -1
echo "full frame TRACE_CTAS=5"; F3D_B200_TRACE_CTAS=5 python tools/ab_bench.py 2>&1 | tail -1
ncu --metrics gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"k_primary|k_trace|k_accum|k_ascent" -s 24 -c 14 --csv --log-file gpurun_out/r02i_launches_part8.csv python tools/ab_bench.py --part 0/8 --frames 8 --warmup 8 --repeat 1 > /dev/null 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open("gpurun_out/r02i_launches_part8.csv")))
h=[i for i,r in enumerate(rows) if r and r[0]=="ID"][0]
c=rows[h]; ki=c.index("Kernel Name"); mi=c.index("Metric Name"); vi=c.index("Metric Value")
agg={}
for r in rows[h+1:]:
    if len(r)>vi: agg.setdefault(r[ki][:40],{}).setdefault(r[mi],[]).append(float(r[vi].replace(",","")))
for k,v in agg.items():
    print("NCU part 0/8", k, len(v["gpu__time_duration.sum"]), {m: round(sum(x)/len(x),2) for m,x in v.items()})
PY
