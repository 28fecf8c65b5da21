#!/bin/bash
# second A/B pass: GPU suite on the current defaults, then per-kernel times with 1024 / 2048 depth buckets in the per-tile sort
# (dgr_set_tuning bit 28); emit (range scan with its loads in flight) and preprocess_fwd (12 atomics in flight) vs tools/r2_ab2.sh's run
cd "$(dirname "$0")/.." && mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/ab3_suite.log 2>&1; echo "suite rc=$?" | tee -a gpurun_out/ab3_suite.log; tail -4 gpurun_out/ab3_suite.log
NEW=1; FINE=$((1 | 1<<28))
timeout 300 python tools/tune.py --steps 24 --tunings "1,1,$NEW;1,1,$FINE;1,1,$NEW;1,1,$FINE;1,1,$NEW;1,1,$FINE" > gpurun_out/ab3_cfg2.log 2>&1
timeout 300 python tools/tune.py --steps 12 --opacity init --tunings "1,1,$NEW;1,1,$FINE;1,1,$NEW;1,1,$FINE" > gpurun_out/ab3_cfg2_init.log 2>&1
timeout 300 python tools/tune.py --steps 8 --points 500000 --res 512 --tunings "1,1,$NEW;1,1,$FINE;1,1,$NEW;1,1,$FINE" > gpurun_out/ab3_cfg3.log 2>&1
for f in cfg2 cfg2_init cfg3; do echo "== $f"; grep -E "^1," gpurun_out/ab3_$f.log | cut -c1-300; tail -2 gpurun_out/ab3_$f.log | grep -v "^1," | cut -c1-300; done
