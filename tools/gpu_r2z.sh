#!/bin/bash
# round-2 GPU call Z: ncu --set full of the final library (-> profiles/traffic.json) + the driver's bench line
mkdir -p gpurun_out
ncu --set full --import-source on --clock-control none -k regex:"k_ptrace|k_shade|k_trace|k_accum|k_ascent" -s 14 -c 10 -o gpurun_out/r02z_full -f python tools/ab_bench.py --frames 8 --warmup 8 --repeat 1 > gpurun_out/r02z_full.log 2>&1
tail -1 gpurun_out/r02z_full.log
python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=line 2>&1 | tail -3
