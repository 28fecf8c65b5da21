#!/bin/bash
# A/B of prebuilt library variants (gpurun_scratch/lib*.so) on the bench workload: per-kernel CUDA-event times
cd "$(dirname "$0")/.." && mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
cp dreamgaussian_b200/lib/libdgr_b200.so /tmp/lib_keep.so
for v in A B A B; do
  cp gpurun_scratch/lib$v.so dreamgaussian_b200/lib/libdgr_b200.so
  echo "== variant $v" >> gpurun_out/ab.log
  timeout 300 python tools/tune.py --steps 24 --tunings "1,1,1;1,1,513;1,1,1;1,1,513" >> gpurun_out/ab.log 2>&1
done
cp /tmp/lib_keep.so dreamgaussian_b200/lib/libdgr_b200.so
grep -E "^==|^1," gpurun_out/ab.log | cut -c1-330
