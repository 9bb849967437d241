#!/bin/bash
# round-2 GPU call M: near-field walk of the sun rays (F3D_SUN_NEAR) A/B, TMA-off default, parity, launch list, bench line
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_multigpu.py -m gpu -x -q 2>&1 | tail -3
echo "--- default (sun near walk, k<=4)"
python tools/ab_bench.py 2>&1 | tail -1
echo "--- no near walk"
F3D_B200_LIB=variants/lib_nonear.so python tools/ab_bench.py 2>&1 | tail -1
echo "--- near walk k<=2"
F3D_B200_LIB=variants/lib_near1.so python tools/ab_bench.py 2>&1 | tail -1
echo "--- ascent 2 CTAs/SM (no register cap)"
F3D_B200_LIB=variants/lib_asc2.so python tools/ab_bench.py 2>&1 | tail -1
echo "--- trace CTAs 5 / 6 / 3"
F3D_B200_TRACE_CTAS=5 python tools/ab_bench.py 2>&1 | tail -1
F3D_B200_TRACE_CTAS=6 python tools/ab_bench.py 2>&1 | tail -1
F3D_B200_TRACE_CTAS=3 python tools/ab_bench.py 2>&1 | tail -1
echo "--- fast numerics"
F3D_B200_NUMERICS=fast python tools/ab_bench.py 2>&1 | tail -1
echo "--- 1/8"
python tools/ab_bench.py --part 0/8 2>&1 | tail -1
M=gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__t_sector_hit_rate.pct,sm__warps_active.avg.pct_of_peak_sustained_active
ncu --metrics $M --clock-control none -k regex:"k_ptrace|k_shade|k_trace|k_accum|k_ascent|k_hz" -c 30 --csv --log-file gpurun_out/r02m_launches.csv python tools/ab_bench.py --frames 8 --warmup 8 --repeat 1 > /dev/null 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open("gpurun_out/r02m_launches.csv")))
h=[i for i,r in enumerate(rows) if r and r[0]=="ID"][0]
c=rows[h]; ki=c.index("Kernel Name"); mi=c.index("Metric Name"); vi=c.index("Metric Value")
agg={}
for r in rows[h+1:]:
    if len(r)>vi:
        try: agg.setdefault(r[ki][:34],{}).setdefault(r[mi].split(".")[0][-28:],[]).append(float(r[vi].replace(",","")))
        except ValueError: pass
for k,v in agg.items():
    print("NCU", k, len(list(v.values())[0]), {m: round(sum(x)/len(x),2) for m,x in v.items()})
PY
python bench.py --steps 20 --warmup 5 > gpurun_out/r02m_bench.json 2> gpurun_out/r02m_bench.err
tail -c 600 gpurun_out/r02m_bench.json
