#!/bin/bash
# round-2 GPU call P: per-list k_ascent kernels (instruction-cache fit), sun near walk, escape map A/B
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_multigpu.py -m gpu -x -q --tb=short 2>&1 | tail -15
echo "--- default (sun near walk + escape map)"
F3D_B200_DEBUG=1 python tools/ab_bench.py 2>&1 | grep "escape map\|^AB" | sort | uniq
echo "--- escape map off"
F3D_B200_ESCAPE=0 python tools/ab_bench.py 2>&1 | tail -1
echo "--- fast numerics"
F3D_B200_NUMERICS=fast python tools/ab_bench.py 2>&1 | tail -1
echo "--- 1/8"
python tools/ab_bench.py --part 0/8 2>&1 | tail -1

M=gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__t_sector_hit_rate.pct,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio
ncu --metrics $M --clock-control none -k regex:"k_ptrace|k_shade|k_trace|k_accum|k_ascent|k_hz|k_escape" -c 60 --csv --log-file gpurun_out/r02p_launches.csv python tools/ab_bench.py --frames 8 --warmup 8 --repeat 1 > /dev/null 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open("gpurun_out/r02p_launches.csv")))
h=[i for i,r in enumerate(rows) if r and r[0]=="ID"][0]
c=rows[h]; ki=c.index("Kernel Name"); mi=c.index("Metric Name"); vi=c.index("Metric Value")
agg={}
for r in rows[h+1:]:
    if len(r)>vi:
        try: agg.setdefault(r[ki][:40],{}).setdefault(r[mi].split(".")[0][-26:],[]).append(float(r[vi].replace(",","")))
        except ValueError: pass
for k,v in agg.items():
    print("NCU", k, len(list(v.values())[0]), {m: round(sum(x)/len(x),2) for m,x in v.items()})
PY
