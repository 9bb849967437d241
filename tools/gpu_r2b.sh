#!/bin/bash
# round-2 GPU call B: bottom-up tracer parity + A/B + ncu
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -4
python tools/ab_bench.py 2>&1 | tail -1
for v in td rb16 rb28 lb8 lb12 lb2 p3; do F3D_B200_LIB=variants/lib_$v.so python tools/ab_bench.py 2>&1 | tail -1; done
python tools/ab_bench.py --part 0/8 2>&1 | tail -1
ncu --metrics gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"k_primary|k_trace|k_accum|k_ascent" -s 32 -c 8 --csv --log-file gpurun_out/r02e_launches.csv python tools/ab_bench.py --frames 4 --warmup 8 --repeat 1 > /dev/null 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open("gpurun_out/r02e_launches.csv")))
h=[i for i,r in enumerate(rows) if r and r[0]=="ID"][0]
c=rows[h]; ki=c.index("Kernel Name"); mi=c.index("Metric Name"); vi=c.index("Metric Value")
agg={}
for r in rows[h+1:]:
    if len(r)>vi: agg.setdefault(r[ki][:40],{}).setdefault(r[mi],[]).append(float(r[vi].replace(",","")))
for k,v in agg.items():
    print("NCU", k, {m: round(sum(x)/len(x),2) for m,x in v.items()})
PY
ncu --set full --clock-control none --import-source on -k regex:"k_primary|k_trace|k_ascent" -s 24 -c 3 -o gpurun_out/r02e_full python tools/ab_bench.py --frames 4 --warmup 8 --repeat 1 > gpurun_out/r02e_full.log 2>&1
compute-sanitizer --tool memcheck python tools/ab_bench.py --frames 2 --warmup 3 --repeat 1 --width 256 --height 144 > gpurun_out/r02e_memcheck.log 2>&1; tail -3 gpurun_out/r02e_memcheck.log
