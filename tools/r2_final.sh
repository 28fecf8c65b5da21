#!/bin/bash
# round-2 final single-GPU pass: suite, smoke, default bench line, reference arm, other workloads, ncu evidence, sanitizers
cd "$(dirname "$0")/.." && mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
rm -f gpurun_out/parity_suite.jsonl
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/final_suite.log 2>&1; echo "suite rc=$?" | tee -a gpurun_out/final_suite.log
tail -5 gpurun_out/final_suite.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/final_smoke.log; tail -2 gpurun_out/final_smoke.log | cut -c1-300
timeout 900 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; tail -c 1200 gpurun_out/final_bench.json; tail -3 gpurun_out/final_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final_bench_reference.json 2> gpurun_out/final_bench_reference.err; tail -c 400 gpurun_out/final_bench_reference.json
timeout 600 python bench.py --workload cfg3 --no-cpu-baseline > gpurun_out/final_bench_cfg3.json 2> gpurun_out/final_bench_cfg3.err; tail -c 700 gpurun_out/final_bench_cfg3.json
timeout 600 python bench.py --workload cfg5 --no-cpu-baseline > gpurun_out/final_bench_cfg5.json 2> gpurun_out/final_bench_cfg5.err; tail -c 700 gpurun_out/final_bench_cfg5.json
timeout 400 python tools/tune.py --tunings "1,1,1;1,1,513;1,2,1" > gpurun_out/final_tune_trained.log 2>&1; tail -4 gpurun_out/final_tune_trained.log
timeout 400 python tools/tune.py --opacity init --tunings "1,1,1;1,2,1" > gpurun_out/final_tune_init.log 2>&1; tail -3 gpurun_out/final_tune_init.log
sed -i 's/r[0-9]_launches/final_launches/g; s/r[0-9]_prof/final_prof/g' tools/r2_profile.sh; bash tools/r2_profile.sh
sed -i 's/r2_sanitizer/final_sanitizer/g' tools/r2_sanitize.sh; bash tools/r2_sanitize.sh
