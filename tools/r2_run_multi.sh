#!/bin/bash
# round-2 multi-GPU check (gpurun --gpus N): own all-reduce kernels (in-kernel barriers, unrolled) vs NCCL, then the bench line
N=${1:-2}
cd "$(dirname "$0")/.." && mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests/test_parity_gpu.py -q -k "allreduce" > gpurun_out/r2_multi_test.log 2>&1; echo "allreduce test rc=$?" | tee -a gpurun_out/r2_multi_test.log; tail -3 gpurun_out/r2_multi_test.log
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 300 --warmup 10 --no-rows \
      > gpurun_out/r2_bench_n${N}_$name.json 2> gpurun_out/r2_bench_n${N}_$name.err
  echo "$name rc=$?"; python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2_bench_n${N}_$name.json").read().strip().splitlines()[-1])
    print("$name", "ms/step", round(d["ms_per_step"], 4), "value", "%.3e" % d["value"], d["config"]["parallelism"], d["config"]["collective_check"], "e2e", d["e2e"] and round(d["e2e"]["ms_per_step"], 3))
except Exception as e:
    print("$name", "no json:", e); print(open("gpurun_out/r2_bench_n${N}_$name.err").read()[-1500:])
PY
}
run default DGR_X=0
run inkernel DGR_INKERNEL_BARRIERS=1
run p2p DGR_NO_MULTIMEM=1
run nccl DGR_NO_PEER=1
timeout 300 python bench.py --steps 300 --warmup 10 --no-rows --no-e2e --no-cpu-baseline > gpurun_out/r2_bench_n1_ref.json 2> gpurun_out/r2_bench_n1_ref.err
python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_n1_ref.json').read().strip().splitlines()[-1]); print('N=1 ms/step', round(d['ms_per_step'],4))"
