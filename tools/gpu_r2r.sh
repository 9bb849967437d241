#!/bin/bash
# round-2 GPU call R (8 GPUs): the driver's scaling launch line at N = 8 (C2 headline + C5 secondary), then N = 4 and N = 2; multi-GPU parity tests
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 600 python -m pytest tests/test_multigpu.py -m gpu -x -q --tb=short 2>&1 | tail -6
for n in 8 4 2; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/r02r_bench_n$n.json 2> gpurun_out/r02r_bench_n$n.err
  python - <<PY
import json
for line in open("gpurun_out/r02r_bench_n$n.json"):
    if line.startswith("{"):
        d=json.loads(line)
        print("N=$n", "value", round(d["value"],1), "ms/step", round(d["ms_per_step"],4), "e2e", d["e2e"] and round(d["e2e"]["value"],1), "call_ms", d["e2e"] and round(d["e2e"]["call_ms"],1), "c5", d.get("secondary_c5") and round(d["secondary_c5"].get("value",0),1), "identical", d.get("bit_identical_to_1gpu"))
PY
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 8 --steps 256 --warmup 8 --no-secondary --no-identity > gpurun_out/r02r_bench_n8_256.json 2> gpurun_out/r02r_bench_n8_256.err
tail -c 300 gpurun_out/r02r_bench_n8_256.json
grep -c "NCCL INFO" gpurun_out/r02r_bench_n8.err; grep "nranks\|NVLS" gpurun_out/r02r_bench_n8.err | head -5
