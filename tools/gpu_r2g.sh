#!/bin/bash
# round-2 GPU call G: the driver's round-end sequence on one GPU: full -m gpu suite, smoke(), both bench arms
mkdir -p gpurun_out
(time python -m pytest tests -m gpu -x -q 2>&1 | tail -6) 2>&1 | tail -10
(time python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3) 2>&1 | tail -6
(time python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02g_ref.json 2> gpurun_out/r02g_ref.err) 2>&1 | tail -3
(time python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02g_bench.json 2> gpurun_out/r02g_bench.err) 2>&1 | tail -3
tail -3 gpurun_out/r02g_bench.err
python - <<'PY'
import json
for f in ("gpurun_out/r02g_ref.json", "gpurun_out/r02g_bench.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "unparsable", e); continue
    print(f, "value %.1f ms/step %.4f" % (d["value"], d["ms_per_step"]), "frac", (d.get("roofline") or {}).get("frac"))
    print("  e2e", d.get("e2e"))
    print("  secondary", d.get("secondary"))
    print("  c5", d.get("secondary_c5"))
    print("  clocks", d.get("clocks"), "launches", d.get("gpu_launches"))
    print("  cpu", d.get("cpu_baseline"))
    for k, v in (d.get("widened_rows") or {}).items():
        print("  row", k, {kk: vv for kk, vv in v.items() if kk in ("value", "unit", "error", "wall_s", "metric", "e2e")})
PY
(time python bench.py --gpus 1 --steps 256 --warmup 8 --no-rows --no-cpu-baseline --no-secondary > gpurun_out/r02g_bench256.json 2>/dev/null) 2>&1 | tail -3
python -c "
import json; d=json.loads(open('gpurun_out/r02g_bench256.json').read().strip().splitlines()[-1]); print('256: value %.1f ms/step %.4f frac %.4f e2e %.1f call_ms %.1f readback %.2f' % (d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['e2e']['call_ms'], d['e2e']['readback_ms']))"
