#!/bin/bash
# 2-GPU check of the final library (gpurun --gpus 2): the two multi-GPU tests the 1-GPU suite skips (own all-reduce vs NCCL; the
# reduce-scatter fused into the per-Gaussian backward, which now stages its inputs by bulk TMA) and the N=2 bench line
cd "$(dirname "$0")/.." && mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 400 python -m pytest tests/test_parity_gpu.py -q -k "allreduce or push" > gpurun_out/multi2_test.log 2>&1; echo "2-GPU tests rc=$?" | tee -a gpurun_out/multi2_test.log; tail -3 gpurun_out/multi2_test.log
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 300 --warmup 10 --no-rows \
    > gpurun_out/multi2_bench_n2.json 2> gpurun_out/multi2_bench_n2.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/multi2_bench_n2.json").read().strip().splitlines()[-1])
    print("N=2 ms/step", round(d["ms_per_step"], 4), "value %.3e" % d["value"], d["config"]["parallelism"], d["config"]["collective_check"], "e2e", d["e2e"] and round(d["e2e"]["ms_per_step"], 3))
except Exception as e:
    print("no json:", e); print(open("gpurun_out/multi2_bench_n2.err").read()[-1500:])
PY
