#!/bin/bash
# round-2 GPU call Z: ncu --set full of the final library (F3D_B200_RAMP=0: full batches of 4 from the first launch, so that every captured launch is a steady-state one) (-> profiles/traffic.json) + the driver's bench line
mkdir -p gpurun_out
F3D_B200_RAMP=0 ncu --set full --import-source on --clock-control none -k regex:"k_ptrace|k_shade|k_trace|k_accum|k_ascent" -s 14 -c 10 -o gpurun_out/r02zz_full -f python tools/ab_bench.py --frames 8 --warmup 8 --repeat 1 > gpurun_out/r02zz_full.log 2>&1
tail -1 gpurun_out/r02zz_full.log
