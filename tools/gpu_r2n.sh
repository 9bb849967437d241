#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "throughput_numerics" --tb=short 2>&1 | tail -30
ncu --set full --import-source on --clock-control none -k regex:"k_ascent" -s 2 -c 1 -o gpurun_out/r02n_ascent -f python tools/ab_bench.py --frames 8 --warmup 8 --repeat 1 > gpurun_out/r02n.log 2>&1
tail -2 gpurun_out/r02n.log
F3D_B200_NUMERICS=fast ncu --set full --import-source on --clock-control none -k regex:"k_ascent" -s 2 -c 1 -o gpurun_out/r02n_ascent_fast -f python tools/ab_bench.py --frames 8 --warmup 8 --repeat 1 > gpurun_out/r02n_fast.log 2>&1
tail -2 gpurun_out/r02n_fast.log
