#!/bin/bash
# round-2 GPU check 5: dynamic queue again (no look-ahead), two-ended vs heavy-end-only, sort without idle key slots
cd "$(dirname "$0")/.." && mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests/test_parity_gpu.py -x -q -k "tiny or small_deg3 or odd_size or big_tiles or mid_20k or cfg1" > gpurun_out/r5_quick.log 2>&1; echo "quick rc=$?" | tee -a gpurun_out/r5_quick.log
tail -3 gpurun_out/r5_quick.log
if grep -q "rc=124" gpurun_out/r5_quick.log; then exit 1; fi
timeout 600 python tools/tune.py --tunings "1,2,1;1,2,1048577;1,1,1;1,1,1048577;1,2,17;1,2,1048593" > gpurun_out/r5_tune_trained.log 2>&1; tail -7 gpurun_out/r5_tune_trained.log
timeout 400 python tools/tune.py --opacity init --tunings "1,2,1;1,2,1048577;1,1,1" > gpurun_out/r5_tune_init.log 2>&1; tail -4 gpurun_out/r5_tune_init.log
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r5_suite.log 2>&1; echo "suite rc=$?" | tee -a gpurun_out/r5_suite.log
tail -5 gpurun_out/r5_suite.log
