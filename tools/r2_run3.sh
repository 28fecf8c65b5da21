#!/bin/bash
# round-2 GPU check 3: one-wave per-Gaussian kernels, coalesced emit prologue, persistent register-key sort, render grid sweep
cd "$(dirname "$0")/.." && mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
rm -f gpurun_out/parity_suite.jsonl
timeout 240 python -m pytest tests/test_parity_gpu.py -x -q -k "tiny or small_deg3 or odd_size or big_tiles" > gpurun_out/r3_quick.log 2>&1; echo "quick rc=$?" | tee -a gpurun_out/r3_quick.log
tail -3 gpurun_out/r3_quick.log
if grep -q "rc=124" gpurun_out/r3_quick.log; then exit 1; fi
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r3_suite.log 2>&1; echo "suite rc=$?" | tee -a gpurun_out/r3_suite.log
tail -12 gpurun_out/r3_suite.log
timeout 600 python tools/tune.py --tunings "1,2,1;1,2,8193;1,2,12289;1,2,16385;1,2,24577;1,2,16401;1,2,131073;1,2,196609;1,2,262145;1,1,1;1,1,262145;1,2,513" > gpurun_out/r3_tune_trained.log 2>&1; tail -13 gpurun_out/r3_tune_trained.log
timeout 400 python tools/tune.py --opacity init --tunings "1,2,1;1,2,16385;1,2,196609;1,1,1" > gpurun_out/r3_tune_init.log 2>&1; tail -5 gpurun_out/r3_tune_init.log
timeout 400 python tools/tune.py --points 2000000 --res 1600 --steps 5 --tunings "1,2,1" > gpurun_out/r3_tune_cfg5.log 2>&1; tail -2 gpurun_out/r3_tune_cfg5.log
sed -i 's/r2_launches/r3_launches/g; s/r2_prof/r3_prof/g' tools/r2_profile.sh; bash tools/r2_profile.sh
timeout 900 python bench.py > gpurun_out/r3_bench_default.json 2> gpurun_out/r3_bench_default.err; tail -c 2500 gpurun_out/r3_bench_default.json; tail -3 gpurun_out/r3_bench_default.err
