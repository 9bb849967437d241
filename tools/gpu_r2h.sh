#!/bin/bash
# round-2 GPU call H (2 GPUs): NCCL parity tests + bench at N=2
mkdir -p gpurun_out
nvidia-smi -L
(time python -m pytest tests/test_multigpu.py -m gpu -x -q 2>&1 | tail -6) 2>&1 | tail -8 | tee gpurun_out/r02h_multigpu_tests.log
(time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02h_bench2.json 2> gpurun_out/r02h_bench2.err) 2>&1 | tail -3
grep -c "NCCL INFO" gpurun_out/r02h_bench2.err; grep -m3 "nranks\|Init COMPLETE" gpurun_out/r02h_bench2.err | cut -c1-200
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02h_bench2.json").read().strip().splitlines()[-1])
print("N=2 value %.1f ms/step %.4f frac %.4f identical %s" % (d["value"], d["ms_per_step"], d["roofline"]["frac"], d.get("bit_identical_to_1gpu")))
print("  e2e", d.get("e2e"))
print("  c5", d.get("secondary_c5"))
PY
(time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 256 --warmup 8 --no-secondary > gpurun_out/r02h_bench2_256.json 2>/dev/null) 2>&1 | tail -3
python -c "
import json; d=json.loads(open('gpurun_out/r02h_bench2_256.json').read().strip().splitlines()[-1]); print('N=2 256: value %.1f ms/step %.4f e2e %.1f call_ms %.1f identical %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['call_ms'], d.get('bit_identical_to_1gpu')))"
