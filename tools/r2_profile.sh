#!/bin/bash
# round-2 profiling pass (one GPU): ncu launch list of a short bench run + one `--set full` capture of every kernel of the step,
# clocks sampled alongside.  Read here afterwards with tools/ncu_kernels_json.py / tools/ncu_summary.py.
cd "$(dirname "$0")/.." && mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
BENCH="python bench.py --no-e2e --no-cpu-baseline --no-rows"
ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 160 --csv --log-file gpurun_out/r2_launches.csv $BENCH --steps 16 --warmup 4 > gpurun_out/r2_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"render_|preprocess_|emit_|tile_" -s 16 -c 16 -f -o gpurun_out/r2_prof $BENCH --steps 3 --warmup 3 > gpurun_out/r2_prof.log 2>&1
ls -la gpurun_out/r2_prof.ncu-rep gpurun_out/r2_launches.csv
tail -3 gpurun_out/r2_prof.log
