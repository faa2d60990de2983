#!/bin/bash
# Round-2 profiling pass (run on the GPU box through gpurun; outputs under gpurun_out/, summaries copied to profiles/).
#   1. launch list of the bench command (per-kernel share of the step)
#   2. ncu --set full of the dominant kernel (one-product seeded sweep) and of the int8 digit-plane GEMM
#   3. compute-sanitizer memcheck + racecheck of the smoke shape
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/r2_launches_bench.json 2> gpurun_out/r2_launches.err
GTB_TC_LIST=32 GTB_TC_QTILES=2 ncu --set full --clock-control none --import-source on -k regex:search_tc_kernel -s 3 -c 1 \
  -o gpurun_out/r2_prof_tch1 -f python scripts/exp_search.py --dtype 3 --reps 1 > gpurun_out/r2_prof_tch1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_i8_kernel -s 2 -c 1 \
  -o gpurun_out/r2_prof_gemm -f python scripts/run_gemm_row.py --t 2 > gpurun_out/r2_prof_gemm.log 2>&1
timeout 900 compute-sanitizer --tool memcheck --log-file gpurun_out/r2_sanitizer_memcheck.log \
  python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_sanitizer_memcheck.out 2>&1
echo "memcheck rc=$?" >> gpurun_out/r2_sanitizer_memcheck.out
timeout 900 compute-sanitizer --tool racecheck --log-file gpurun_out/r2_sanitizer_racecheck.log \
  python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_sanitizer_racecheck.out 2>&1
echo "racecheck rc=$?" >> gpurun_out/r2_sanitizer_racecheck.out
tail -3 gpurun_out/r2_sanitizer_*.log gpurun_out/r2_sanitizer_*.out
