#!/bin/bash
# round-2 GPU call X (8 GPUs): the driver's scaling line at N = 8 with the warmed e2e path, steady-state runs and three run-time knobs
mkdir -p gpurun_out
run() { # tag N steps extra-args... (env passes through)
  tag=$1; n=$2; st=$3; shift 3
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus $n --steps $st --warmup 5 "$@" > gpurun_out/r02x_$tag.json 2> gpurun_out/r02x_$tag.err
  python - <<PY
import json
for line in open("gpurun_out/r02x_$tag.json"):
    if line.startswith("{"):
        d=json.loads(line); e=d.get("e2e")
        print("$tag", "N=$n steps=$st", "Mrays/s", round(d["value"],1), "ms/step", round(d["ms_per_step"],4), "e2e", e and round(e["value"],1), "call_ms", e and round(e["call_ms"],1), "c5", d.get("secondary_c5") and round(d["secondary_c5"].get("value",0),1), "identical", d.get("bit_identical_to_1gpu"), e and e.get("rank0_phases_ms"))
PY
}
run n8_s20 8 20
run n8_s256 8 256 --no-secondary --no-identity
FAST="--no-secondary --no-identity --no-e2e --no-cpu-baseline"
run n8_s256_b 8 256 $FAST
F3D_B200_SETS=3 run n8_s256_sets3 8 256 $FAST
F3D_B200_PRIM_PRIORITY=1 run n8_s256_prio 8 256 $FAST
F3D_B200_BATCH=4 run n8_s256_batch4 8 256 $FAST
run n4_s256 4 256 --no-secondary --no-identity
