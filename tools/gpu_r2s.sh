#!/bin/bash
# round-2 GPU call S (2 GPUs): whole -m gpu suite, viewshed side bench + lanes, tuning sweep, primary-stream priority A/B at N = 2
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --tb=line 2>&1 | tail -8
python tools/bench_viewshed.py > gpurun_out/r02s_viewshed.json 2>&1; tail -c 700 gpurun_out/r02s_viewshed.json; echo
ncu --metrics gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active -k regex:"k_viewshed|k_shadow_mask" -c 2 --csv --log-file gpurun_out/r02s_viewshed_launches.csv python tools/bench_viewshed.py > /dev/null 2>&1
grep "k_viewshed\|k_shadow" gpurun_out/r02s_viewshed_launches.csv | cut -d, -f5,13-16 | tr -d '"' | head -12
echo "--- sweep (C2, 1 GPU)"
python tools/ab_bench.py 2>&1 | tail -1
for v in refill12 refill20 refill24; do F3D_B200_LIB=variants/lib_$v.so python tools/ab_bench.py 2>&1 | tail -1; done
echo "ascent ctas 6 / 12 / 16"; for c in 6 12 16; do F3D_B200_ASCENT_CTAS=$c python tools/ab_bench.py 2>&1 | tail -1; done
echo "batch 8 / sets 3"; F3D_B200_BATCH=8 python tools/ab_bench.py 2>&1 | tail -1; F3D_B200_SETS=3 python tools/ab_bench.py 2>&1 | tail -1
echo "prim priority"; F3D_B200_PRIM_PRIORITY=1 python tools/ab_bench.py 2>&1 | tail -1
echo "part 0/8, prio 0 / 1"; python tools/ab_bench.py --part 0/8 2>&1 | tail -1; F3D_B200_PRIM_PRIORITY=1 python tools/ab_bench.py --part 0/8 2>&1 | tail -1
echo "part 0/2"; python tools/ab_bench.py --part 0/2 2>&1 | tail -1
python tools/gather_only_check.py --scene c2 --world 8 --frames 64 2>&1 | tail -1
echo "--- N=2"
for prio in 0 1; do for st in 20 256; do
  F3D_B200_PRIM_PRIORITY=$prio timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps $st --warmup 8 --no-secondary --no-identity > gpurun_out/r02s_n2_p${prio}_$st.json 2> gpurun_out/r02s_n2_p${prio}_$st.err
  python - <<PY
import json
for line in open("gpurun_out/r02s_n2_p${prio}_$st.json"):
    if line.startswith("{"):
        d=json.loads(line); print("N=2 prio $prio steps $st: ms/step", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"],1), "call_ms", round(d["e2e"]["call_ms"],1), d["e2e"].get("rank0_phases_ms"))
PY
done; done
