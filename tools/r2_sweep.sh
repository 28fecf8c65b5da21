#!/bin/bash
# in-pipeline sweep of the render-kernel tuning word (step_us is the figure of merit; kernels_us are timed one by one)
cd "$(dirname "$0")/.." && mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
for rep in 1 2; do
  timeout 500 python tools/tune.py --steps 24 --tunings "1,1,1;1,1,513;1,1,17;1,1,529;1,1,262657;1,1,328193;1,1,393729;1,1,327681;1,1,393217;1,1,458753;1,1,20993;1,1,25089;1,1,29185;2,1,513;1,2,1;1,2,513;2,2,513;1,1,521;1,1,9" >> gpurun_out/sweep.log 2>&1
done
grep -E "^[0-9]," gpurun_out/sweep.log | python -c "
import sys, json
for l in sys.stdin:
    k, j = l.split(' ', 1); d = json.loads(j); ku = d['kernels_us']
    print('%-14s step %6.1f  fwd %5.1f bwd %5.1f sum %6.1f' % (k, d['step_us'], ku['render_fwd'], ku['render_bwd'], d['sum_us']))"
