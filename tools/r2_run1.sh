#!/bin/bash
# round-2 GPU check 1: persistent render kernels + matrix-free binning + PDL chain + densify kernels — quick parity subset
# (hang guard), full suite, tuning sweep (incl. PDL on/off), parity report, bench smoke; the render-only tree
# (gpurun_scratch/v1) is the fallback that isolates failures
cd "$(dirname "$0")/.." && mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
rm -f gpurun_out/parity_suite.jsonl
timeout 240 python -m pytest tests/test_parity_gpu.py -x -q -k "tiny or small_deg3 or odd_size" > gpurun_out/r2_quick.log 2>&1; echo "quick rc=$?" | tee -a gpurun_out/r2_quick.log
tail -5 gpurun_out/r2_quick.log
if ! grep -q "rc=124" gpurun_out/r2_quick.log; then
  timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r2_suite.log 2>&1; echo "suite rc=$?" | tee -a gpurun_out/r2_suite.log
  tail -40 gpurun_out/r2_suite.log
  timeout 300 python tools/tune.py --tunings "1,2,1;1,2,9;2,2,1;1,1,1;2,1,1;4,2,1" > gpurun_out/r2_tune_trained.log 2>&1; tail -7 gpurun_out/r2_tune_trained.log
  timeout 300 python tools/tune.py --opacity init --tunings "1,2,1;2,2,1;1,1,1;2,1,1" > gpurun_out/r2_tune_init.log 2>&1; tail -5 gpurun_out/r2_tune_init.log
  timeout 900 python tools/parity_report.py > gpurun_out/r2_parity_report.log 2>&1; tail -15 gpurun_out/r2_parity_report.log
  timeout 600 python bench.py --steps 100 --warmup 5 > gpurun_out/r2_bench_cfg2.json 2> gpurun_out/r2_bench_cfg2.err; tail -c 3000 gpurun_out/r2_bench_cfg2.json; tail -5 gpurun_out/r2_bench_cfg2.err
  timeout 600 python bench.py --workload cfg5 --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2_bench_cfg5.json 2> gpurun_out/r2_bench_cfg5.err; tail -c 1500 gpurun_out/r2_bench_cfg5.json; tail -5 gpurun_out/r2_bench_cfg5.err
fi
if ! grep -q " passed" gpurun_out/r2_suite.log 2>/dev/null || grep -q "failed" gpurun_out/r2_suite.log; then
  cd gpurun_scratch/v1
  timeout 240 python -m pytest tests/test_parity_gpu.py -x -q -k "tiny or small_deg3 or odd_size" > ../../gpurun_out/r2_v1_quick.log 2>&1; echo "v1 quick rc=$?" | tee -a ../../gpurun_out/r2_v1_quick.log
  tail -5 ../../gpurun_out/r2_v1_quick.log
  if ! grep -q "rc=124" ../../gpurun_out/r2_v1_quick.log; then
    timeout 900 python -m pytest tests -m gpu -q > ../../gpurun_out/r2_v1_suite.log 2>&1; echo "v1 suite rc=$?" | tee -a ../../gpurun_out/r2_v1_suite.log
    tail -25 ../../gpurun_out/r2_v1_suite.log
    timeout 300 python tools/tune.py --tunings "1,2,1;2,2,1" > ../../gpurun_out/r2_v1_tune_trained.log 2>&1; tail -3 ../../gpurun_out/r2_v1_tune_trained.log
  fi
  cd ../..
fi
