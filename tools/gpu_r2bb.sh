#!/bin/bash
# final 2-GPU check: multi-GPU parity tests + the driver's line at N = 2 with the shipped library
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_multigpu.py -m gpu -q --tb=line 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02bb_n2.json 2> gpurun_out/r02bb_n2.err
python - <<'PY'
import json
d=[json.loads(l) for l in open("gpurun_out/r02bb_n2.json") if l.startswith("{")][-1]; e=d["e2e"]
print("N=2 steps 20: Mrays/s", round(d["value"],1), "ms/step", round(d["ms_per_step"],4), "frac", round(d["roofline"]["frac"],3), "e2e", round(e["value"],1), "call_ms", round(e["call_ms"],1), "identical", d.get("bit_identical_to_1gpu"), "c5", d["secondary_c5"] and round(d["secondary_c5"]["value"],1), e.get("rank0_phases_ms"))
PY
grep -c "NCCL INFO" gpurun_out/r02bb_n2.err
