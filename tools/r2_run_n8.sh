#!/bin/bash
# round-2 8-GPU check (gpurun --gpus 8): phase breakdown of the multi-GPU step for every collective variant, then the bench line at N=8 and N=4
cd "$(dirname "$0")/.." && mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 tools/allreduce_phases.py --steps 80 > gpurun_out/phases_n8.json 2> gpurun_out/phases_n8.err
echo "phases rc=$?"; grep -v "NCCL version" gpurun_out/phases_n8.json | python -c "
import json,sys
d=json.load(sys.stdin)
for k,v in d.items():
    print(k, v if not isinstance(v,dict) else {a:b for a,b in v.items()})
"; tail -3 gpurun_out/phases_n8.err | cut -c1-300
for N in 8 4; do
  timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus $N --steps 300 --warmup 10 --no-rows --no-cpu-baseline \
      > gpurun_out/r2_bench_n${N}.json 2> gpurun_out/r2_bench_n${N}.err
  echo "bench N=$N rc=$?"; python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2_bench_n${N}.json").read().strip().splitlines()[-1])
    print("N=${N}", "ms/step", round(d["ms_per_step"], 4), "value", "%.3e" % d["value"], d["config"]["parallelism"], d["config"]["collective_check"], "e2e", d["e2e"] and round(d["e2e"]["ms_per_step"], 3))
except Exception as e:
    print("N=${N}", "no json:", e); print(open("gpurun_out/r2_bench_n${N}.err").read()[-1500:])
PY
done
