#!/bin/bash
# round-2 GPU call A: parity of the new traversal on the device, A/B of the build variants, ncu of the default build
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader
python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -4
python tools/ab_bench.py 2>&1 | tail -1
for v in r1 cull p3 t5 wd16 rb28 rb20; do F3D_B200_LIB=variants/lib_$v.so python tools/ab_bench.py 2>&1 | tail -1; done
python tools/ab_bench.py --part 0/8 2>&1 | tail -1
python tools/ab_bench.py --spp 8 --frames 16 2>&1 | tail -1
ncu --metrics gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"k_primary|k_trace|k_accum" -s 24 -c 9 --csv --log-file gpurun_out/r02a_launches.csv python tools/ab_bench.py --frames 4 --warmup 8 --repeat 1 > /dev/null 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open("gpurun_out/r02a_launches.csv")))
h=[i for i,r in enumerate(rows) if r and r[0]=="ID"][0]
c=rows[h]; ki=c.index("Kernel Name"); mi=c.index("Metric Name"); vi=c.index("Metric Value")
agg={}
for r in rows[h+1:]:
    if len(r)>vi: agg.setdefault(r[ki][:40],{}).setdefault(r[mi],[]).append(float(r[vi].replace(",","")))
for k,v in agg.items():
    print("NCU", k, {m: round(sum(x)/len(x),2) for m,x in v.items()})
PY
ncu --set full --clock-control none --import-source on -k regex:"k_primary|k_trace" -s 16 -c 2 -o gpurun_out/r02a_full python tools/ab_bench.py --frames 4 --warmup 8 --repeat 1 > gpurun_out/r02a_full.log 2>&1
ls -la gpurun_out/
