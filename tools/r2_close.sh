#!/bin/bash
# the round's last look at one B200: suite, smoke, default bench line (with roofline.traffic / issue and the e2e timeline), reference arm
cd "$(dirname "$0")/.." && mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/close_suite.log 2>&1; echo "suite rc=$?" | tee -a gpurun_out/close_suite.log; tail -3 gpurun_out/close_suite.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/close_smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/close_smoke.log
timeout 900 python bench.py > gpurun_out/close_bench.json 2> gpurun_out/close_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/close_bench.json").read().strip().splitlines()[-1])
print("ms/step", d["ms_per_step"], "value %.4g" % d["value"], "traffic", d["roofline"]["traffic"], "e2e", d["e2e"]["ms_per_step"], "timeline", bool(d["e2e"]["timeline"]))
PY
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/close_bench_reference.json 2> gpurun_out/close_bench_reference.err; tail -c 200 gpurun_out/close_bench_reference.json
