#!/bin/bash
# A/B of the bulk-TMA input staging of the per-Gaussian kernels (dgr_set_tuning bits 24 / 25) and of the per-row step 2 of the
# backward render (bit 26) on one B200: the GPU suite first (new defaults), then per-kernel CUDA-event times per variant.
cd "$(dirname "$0")/.." && mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/ab2_suite.log 2>&1; echo "suite rc=$?" | tee -a gpurun_out/ab2_suite.log; tail -4 gpurun_out/ab2_suite.log
NEW=1; OLDF=$((1 | 1<<24)); OLDB=$((1 | 1<<25)); OLDR=$((1 | 1<<26)); OLD=$((1 | 7<<24)); CAP5=$((1 | 5<<16))
T="1,1,$OLD;1,1,$NEW;1,1,$OLDF;1,1,$OLDB;1,1,$OLDR;1,1,$CAP5;1,1,$OLD;1,1,$NEW"
timeout 300 python tools/tune.py --steps 24 --tunings "$T" > gpurun_out/ab2_cfg2.log 2>&1
timeout 300 python tools/tune.py --steps 12 --opacity init --tunings "1,1,$OLD;1,1,$NEW;1,1,$OLDR;1,1,$OLD;1,1,$NEW" > gpurun_out/ab2_cfg2_init.log 2>&1
timeout 300 python tools/tune.py --steps 8 --points 500000 --res 512 --tunings "1,1,$OLD;1,1,$NEW;1,1,$OLDF;1,1,$OLDB;1,1,$OLD;1,1,$NEW" > gpurun_out/ab2_cfg3.log 2>&1
timeout 400 python tools/tune.py --steps 6 --points 2000000 --res 1600 --tunings "1,1,$OLD;1,1,$NEW;1,1,$OLDF;1,1,$OLDB;1,1,$OLD;1,1,$NEW" > gpurun_out/ab2_cfg5.log 2>&1
for f in cfg2 cfg2_init cfg3 cfg5; do echo "== $f"; grep -E "^1," gpurun_out/ab2_$f.log | cut -c1-300; tail -2 gpurun_out/ab2_$f.log | grep -v "^1," | cut -c1-300; done
