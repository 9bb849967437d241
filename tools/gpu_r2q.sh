#!/bin/bash
# round-2 GPU call Q: whole -m gpu suite, bench lines (c2 default, c4, reference arm), sanitizer passes, ncu --set full of the frame kernels
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q --tb=short 2>&1 | tail -8
python bench.py > gpurun_out/r02q_bench.json 2> gpurun_out/r02q_bench.err; tail -c 400 gpurun_out/r02q_bench.json; echo
python bench.py --steps 256 --warmup 8 --no-secondary --no-cpu-baseline > gpurun_out/r02q_bench256.json 2>> gpurun_out/r02q_bench.err
python bench.py --config c4 --steps 128 --warmup 4 --no-cpu-baseline --no-e2e > gpurun_out/r02q_bench_c4.json 2>> gpurun_out/r02q_bench.err; tail -c 900 gpurun_out/r02q_bench_c4.json; echo
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02q_ref.json 2>> gpurun_out/r02q_bench.err
for tool in memcheck racecheck initcheck; do
  timeout 600 compute-sanitizer --tool $tool python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02q_sanitizer_$tool.log 2>&1
  echo "sanitizer $tool: $(grep -c 'ERROR SUMMARY: 0 errors' gpurun_out/r02q_sanitizer_$tool.log) clean; $(grep 'ERROR SUMMARY\|RACECHECK SUMMARY' gpurun_out/r02q_sanitizer_$tool.log | tail -1)"
done
ncu --set full --import-source on --clock-control none -k regex:"k_ptrace|k_shade|k_trace|k_accum|k_ascent" -s 14 -c 10 -o gpurun_out/r02q_full -f python tools/ab_bench.py --frames 8 --warmup 8 --repeat 1 > gpurun_out/r02q_full.log 2>&1
tail -2 gpurun_out/r02q_full.log
