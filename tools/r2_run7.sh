#!/bin/bash
# round-2 GPU check 7: records staged by id (bulk TMA per record) vs the sorted copy, per shape
cd "$(dirname "$0")/.." && mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests/test_parity_gpu.py -x -q -k "tiny or small_deg3 or odd_size or big_tiles or huge_tiles or mid_20k or cfg1 or capacity or twice or accumulation" > gpurun_out/r7_quick.log 2>&1; echo "quick rc=$?" | tee -a gpurun_out/r7_quick.log
tail -3 gpurun_out/r7_quick.log
if grep -q "rc=124" gpurun_out/r7_quick.log; then exit 1; fi
timeout 600 python tools/tune.py --tunings "1,1,1;1,1,8388609;1,2,8388609" > gpurun_out/r7_tune_trained.log 2>&1; tail -4 gpurun_out/r7_tune_trained.log
timeout 400 python tools/tune.py --opacity init --tunings "1,1,1;1,1,8388609;1,2,1;1,2,8388609" > gpurun_out/r7_tune_init.log 2>&1; tail -5 gpurun_out/r7_tune_init.log
timeout 400 python tools/tune.py --points 2000000 --res 1600 --steps 5 --tunings "1,1,1;1,1,4194305;1,2,1" > gpurun_out/r7_tune_cfg5.log 2>&1; tail -4 gpurun_out/r7_tune_cfg5.log
timeout 400 python tools/tune.py --points 500000 --res 512 --steps 5 --tunings "1,1,1;1,1,4194305;1,2,1" > gpurun_out/r7_tune_cfg3.log 2>&1; tail -4 gpurun_out/r7_tune_cfg3.log
DGR_TUNING=1,1,8388609 timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r7_suite_lazy.log 2>&1; echo "suite(lazy) rc=$?" | tee -a gpurun_out/r7_suite_lazy.log
tail -5 gpurun_out/r7_suite_lazy.log
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r7_suite.log 2>&1; echo "suite rc=$?" | tee -a gpurun_out/r7_suite.log
tail -5 gpurun_out/r7_suite.log
