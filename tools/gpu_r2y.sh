#!/bin/bash
# round-2 GPU call Y (2 GPUs): k_ascent prefetch / 4-CTA A/B; session_create with batches of 8 under the old and the new cache cap
mkdir -p gpurun_out
python tools/ab_bench.py 2>&1 | tail -1
for v in pf asc4 pf_asc4; do F3D_B200_LIB=variants/lib_$v.so python tools/ab_bench.py 2>&1 | tail -1; done
python tools/ab_bench.py 2>&1 | tail -1
F3D_B200_LIB=variants/lib_pf.so python tools/ab_bench.py 2>&1 | tail -1
for cap in 4096 10240; do
  F3D_B200_CACHE_MB=$cap F3D_B200_BATCH=8 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 2 --steps 20 --warmup 5 --no-secondary --no-identity --no-cpu-baseline > gpurun_out/r02y_n2_cap$cap.json 2> gpurun_out/r02y_n2_cap$cap.err
  python - <<PY
import json
for line in open("gpurun_out/r02y_n2_cap$cap.json"):
    if line.startswith("{"):
        d=json.loads(line); e=d["e2e"]; print("N=2 batch 8 cap $cap MB: ms/step", round(d["ms_per_step"],4), "e2e", round(e["value"],1), "call_ms", round(e["call_ms"],1), e.get("rank0_phases_ms"))
PY
done
