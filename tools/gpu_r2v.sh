#!/bin/bash
# round-2 GPU call V: final ncu --set full of the shipped frame kernels (-> profiles/traffic.json), parity suite, default bench line
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --tb=line 2>&1 | tail -4
ncu --set full --import-source on --clock-control none -k regex:"k_ptrace|k_shade|k_trace|k_accum|k_ascent" -s 14 -c 10 -o gpurun_out/r02v_full -f python tools/ab_bench.py --frames 8 --warmup 8 --repeat 1 > gpurun_out/r02v_full.log 2>&1
tail -1 gpurun_out/r02v_full.log
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02v_bench20.json 2> gpurun_out/r02v_bench.err
python bench.py > gpurun_out/r02v_bench256.json 2>> gpurun_out/r02v_bench.err
python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02v_ref.json 2>> gpurun_out/r02v_bench.err
python bench.py --config c3 --steps 8 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/r02v_bench_c3.json 2>> gpurun_out/r02v_bench.err
python bench.py --config c4 --steps 128 --warmup 4 --no-cpu-baseline --no-e2e > gpurun_out/r02v_bench_c4.json 2>> gpurun_out/r02v_bench.err
python - <<'PY'
import json
for f in ("r02v_bench20","r02v_bench256","r02v_ref","r02v_bench_c3","r02v_bench_c4"):
    for line in open(f"gpurun_out/{f}.json"):
        if line.startswith("{"):
            d=json.loads(line); print(f, round(d["value"],1), round(d["ms_per_step"],4), d.get("roofline",{}).get("frac"), d.get("e2e") and round(d["e2e"]["value"],1), d.get("smoke_over") and d["smoke_over"].get("kernel_ms"))
PY
