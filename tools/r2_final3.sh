#!/bin/bash
# last closing pass (render_fwd: deferred cost filing): forward/backward parity subset first, then tools/r2_final2.sh without the suite
cd "$(dirname "$0")/.." && mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 400 python -m pytest tests/test_parity_gpu.py tests/test_reference_caller_gpu.py -x -q -k "not cfg5 and not cfg3 and not initial_opacity and not 1m" > gpurun_out/final3_tests.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/final3_tests.log; tail -3 gpurun_out/final3_tests.log
SKIP_SUITE=1 bash tools/r2_final2.sh
timeout 200 python tools/tune.py --steps 24 --tunings "1,1,1;1,1,1" > gpurun_out/final3_tune.log 2>&1; grep -E "^1," gpurun_out/final3_tune.log | cut -c1-300
