#!/bin/bash
# round-2 GPU call U: parked-leaf primary traversal (F3D_PRIMARY_PARK) A/B
mkdir -p gpurun_out
python tools/ab_bench.py 2>&1 | tail -1
for v in park8 park12 park16 park24; do F3D_B200_LIB=variants/lib_$v.so python tools/ab_bench.py 2>&1 | tail -1; done
python tools/ab_bench.py 2>&1 | tail -1
M=gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active
for v in default park12; do
  if [ $v != default ]; then export F3D_B200_LIB=variants/lib_$v.so; fi
  ncu --metrics $M --clock-control none -k regex:"k_ptrace" -s 2 -c 2 --csv --log-file gpurun_out/r02u_ptrace_$v.csv python tools/ab_bench.py --frames 8 --warmup 8 --repeat 1 > /dev/null 2>&1
  grep k_ptrace gpurun_out/r02u_ptrace_$v.csv | awk -F'","' '{print "'$v'", $(NF-2), $NF}' | tr -d '"' | head -8
done
