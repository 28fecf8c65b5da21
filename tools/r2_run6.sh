#!/bin/bash
# round-2 GPU check 6: backward work ordered by the cost the forward measured
cd "$(dirname "$0")/.." && mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests/test_parity_gpu.py -x -q -k "tiny or small_deg3 or odd_size or big_tiles or mid_20k or cfg1 or capacity or twice or accumulation" > gpurun_out/r6_quick.log 2>&1; echo "quick rc=$?" | tee -a gpurun_out/r6_quick.log
tail -3 gpurun_out/r6_quick.log
if grep -q "rc=124" gpurun_out/r6_quick.log; then exit 1; fi
timeout 600 python tools/tune.py --tunings "1,2,1;1,2,2097153;1,1,1;1,1,2097153;2,2,1" > gpurun_out/r6_tune_trained.log 2>&1; tail -6 gpurun_out/r6_tune_trained.log
timeout 400 python tools/tune.py --opacity init --tunings "1,2,1;1,2,2097153;1,1,1" > gpurun_out/r6_tune_init.log 2>&1; tail -4 gpurun_out/r6_tune_init.log
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r6_suite.log 2>&1; echo "suite rc=$?" | tee -a gpurun_out/r6_suite.log
tail -5 gpurun_out/r6_suite.log
