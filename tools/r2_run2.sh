#!/bin/bash
# round-2 GPU check 2: batched alpha evaluation in the render kernels, aggregated atomics in the tile scan, new parity rule;
# tuning sweep over the batch sizes, ncu launch list + full capture, compute-sanitizer
cd "$(dirname "$0")/.." && mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
rm -f gpurun_out/parity_suite.jsonl
timeout 240 python -m pytest tests/test_parity_gpu.py -x -q -k "tiny or small_deg3 or odd_size" > gpurun_out/r2_quick.log 2>&1; echo "quick rc=$?" | tee -a gpurun_out/r2_quick.log
tail -3 gpurun_out/r2_quick.log
if grep -q "rc=124" gpurun_out/r2_quick.log; then exit 1; fi
[ -x gpurun_scratch/tma_gather_probe ] && { timeout 120 gpurun_scratch/tma_gather_probe > gpurun_out/r2_tma_probe.log 2>&1; cat gpurun_out/r2_tma_probe.log; }
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r2_suite.log 2>&1; echo "suite rc=$?" | tee -a gpurun_out/r2_suite.log
tail -12 gpurun_out/r2_suite.log
timeout 400 python tools/tune.py --tunings "1,2,1;1,2,17;1,2,33;1,2,257;1,2,1025;1,1,1;1,1,1025;2,2,1;1,2,9" > gpurun_out/r2_tune_trained.log 2>&1; tail -10 gpurun_out/r2_tune_trained.log
timeout 400 python tools/tune.py --opacity init --tunings "1,2,1;1,2,33;1,2,257;1,2,1025;1,1,1;1,1,1025" > gpurun_out/r2_tune_init.log 2>&1; tail -7 gpurun_out/r2_tune_init.log
bash tools/r2_profile.sh
bash tools/r2_sanitize.sh
timeout 900 python bench.py > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err; tail -c 2500 gpurun_out/r2_bench_default.json; tail -3 gpurun_out/r2_bench_default.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err; tail -c 600 gpurun_out/r2_bench_reference.json
