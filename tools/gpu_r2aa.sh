#!/bin/bash
# round-2 GPU call AA: early AOV read-back A/B on the driver's line, parity, final ncu capture (full batches) of the shipped library
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_smoke.py tests/test_aether.py -m gpu -q --tb=line 2>&1 | tail -3
for rep in 1 2; do for r in 0 1; do
  F3D_B200_EARLY_AOVS=$r python bench.py --gpus 1 --steps 20 --warmup 5 --no-secondary --no-cpu-baseline --no-rows > gpurun_out/early_$r.json 2> gpurun_out/early.err
  python - <<PY
import json
d=[json.loads(l) for l in open("gpurun_out/early_$r.json") if l.startswith("{")][-1]; e=d["e2e"]
print("early $r: ms/step", round(d["ms_per_step"],4), "e2e", round(e["value"],1), "call_ms", round(e["call_ms"],2), "setup", round(e["setup_ms"],2), "frames", round(e["frames_ms"],2), "readback", round(e["readback_ms"],2))
PY
done; done
F3D_B200_RAMP=0 ncu --set full --import-source on --clock-control none -k regex:"k_ptrace|k_shade|k_trace|k_accum|k_ascent" -s 14 -c 10 -o gpurun_out/r02aa_full -f python tools/ab_bench.py --frames 8 --warmup 8 --repeat 1 > gpurun_out/r02aa_full.log 2>&1
tail -1 gpurun_out/r02aa_full.log
