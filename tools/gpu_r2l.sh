#!/bin/bash
# round-2 GPU call L: sun horizon strips + TMA staging A/B, parity, launch list
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_multigpu.py -m gpu -x -q 2>&1 | tail -3
echo "--- default (horizon + TMA staging)"
python tools/ab_bench.py 2>&1 | tail -1
echo "--- horizon off"
F3D_B200_SUN_HORIZON=0 python tools/ab_bench.py 2>&1 | tail -1
echo "--- staging off (table only)"
F3D_B200_TMA_STAGE=0 python tools/ab_bench.py 2>&1 | tail -1
echo "--- no table, no staging (compile-time)"
F3D_B200_LIB=variants/lib_notma.so python tools/ab_bench.py 2>&1 | tail -1
echo "--- fast numerics"
F3D_B200_NUMERICS=fast python tools/ab_bench.py 2>&1 | tail -1
echo "--- 1/8"
python tools/ab_bench.py --part 0/8 2>&1 | tail -1
F3D_B200_SUN_HORIZON=0 python tools/ab_bench.py --part 0/8 2>&1 | tail -1
F3D_B200_DEBUG=1 python tools/ab_bench.py --frames 8 --repeat 1 2>&1 | grep forge3d_b200 | sort | uniq | head
M=gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__t_sector_hit_rate.pct,smsp__pcsamp_warps_issue_stalled_long_scoreboard
for tag in default notma; do
  if [ $tag = notma ]; then export F3D_B200_LIB=variants/lib_notma.so; fi
  ncu --metrics $M --clock-control none -k regex:"k_ptrace|k_shade|k_trace|k_accum|k_ascent|k_hz" -c 30 --csv --log-file gpurun_out/r02l_launches_$tag.csv python tools/ab_bench.py --frames 8 --warmup 8 --repeat 1 > /dev/null 2>&1
done
unset F3D_B200_LIB
python - <<'PY'
import csv
for tag in ("default", "notma"):
    rows=list(csv.reader(open(f"gpurun_out/r02l_launches_{tag}.csv")))
    h=[i for i,r in enumerate(rows) if r and r[0]=="ID"][0]
    c=rows[h]; ki=c.index("Kernel Name"); mi=c.index("Metric Name"); vi=c.index("Metric Value")
    agg={}
    for r in rows[h+1:]:
        if len(r)>vi:
            try: agg.setdefault(r[ki][:34],{}).setdefault(r[mi].split(".")[0][-28:],[]).append(float(r[vi].replace(",","")))
            except ValueError: pass
    for k,v in agg.items():
        print("NCU", tag, k, len(list(v.values())[0]), {m: round(sum(x)/len(x),2) for m,x in v.items()})
PY
