#!/bin/bash
# round-2 GPU call K: gather-only tests, fast-numerics A/B + tolerance check, ncu --set full of the current frame kernels
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_multigpu.py -m gpu -x -q 2>&1 | tail -3
python tools/ab_bench.py 2>&1 | tail -1
F3D_B200_NUMERICS=fast python tools/ab_bench.py 2>&1 | tail -1
python tools/numerics_check.py --scene c2 --frames 256 2>&1 | tail -1
python tools/numerics_check.py --scene golden --frames 128 2>&1 | tail -1
python tools/gather_only_check.py --scene c2 --world 8 --frames 64 2>&1 | tail -1
python tools/gather_only_check.py --scene c2 --world 8 --frames 64 --block-rows 144 2>&1 | tail -1
ncu --set full --import-source on --clock-control none -k regex:"k_ptrace|k_shade|k_trace|k_accum|k_ascent" -s 10 -c 9 -o gpurun_out/r02k_full -f python tools/ab_bench.py --frames 8 --warmup 8 --repeat 1 > gpurun_out/r02k_full.log 2>&1
tail -2 gpurun_out/r02k_full.log
F3D_B200_NUMERICS=fast ncu --set full --import-source on --clock-control none -k regex:"k_ptrace|k_shade|k_trace|k_accum|k_ascent" -s 10 -c 9 -o gpurun_out/r02k_full_fast -f python tools/ab_bench.py --frames 8 --warmup 8 --repeat 1 > gpurun_out/r02k_full_fast.log 2>&1
tail -2 gpurun_out/r02k_full_fast.log
