#!/bin/bash
# round-2 GPU check 4: render work queue consumed from both ends, next-item prefetch, static empties
cd "$(dirname "$0")/.." && mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
rm -f gpurun_out/parity_suite.jsonl
timeout 240 python -m pytest tests/test_parity_gpu.py -x -q -k "tiny or small_deg3 or odd_size or big_tiles" > gpurun_out/r4_quick.log 2>&1; echo "quick rc=$?" | tee -a gpurun_out/r4_quick.log
tail -3 gpurun_out/r4_quick.log
if grep -q "rc=124" gpurun_out/r4_quick.log; then exit 1; fi
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r4_suite.log 2>&1; echo "suite rc=$?" | tee -a gpurun_out/r4_suite.log
tail -8 gpurun_out/r4_suite.log
timeout 600 python tools/tune.py --tunings "1,2,1;1,2,1048577;1,1,1;1,1,1048577;2,2,1;1,2,17;1,2,1048593" > gpurun_out/r4_tune_trained.log 2>&1; tail -8 gpurun_out/r4_tune_trained.log
timeout 400 python tools/tune.py --opacity init --tunings "1,2,1;1,2,1048577;1,1,1" > gpurun_out/r4_tune_init.log 2>&1; tail -4 gpurun_out/r4_tune_init.log
sed -i 's/r[23]_launches/r4_launches/g; s/r[23]_prof/r4_prof/g' tools/r2_profile.sh; bash tools/r2_profile.sh
timeout 900 python bench.py > gpurun_out/r4_bench_default.json 2> gpurun_out/r4_bench_default.err; tail -c 1500 gpurun_out/r4_bench_default.json; tail -3 gpurun_out/r4_bench_default.err
