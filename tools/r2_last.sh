#!/bin/bash
# round-2 closing pass on one B200 (after the last kernel edit): suite + smoke, ncu launch list + full capture, the capture's
# per-launch numbers written where bench.py looks for them, then the default bench line (now with roofline.traffic / issue)
cd "$(dirname "$0")/.." && mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/last_suite.log 2>&1; echo "suite rc=$?" | tee -a gpurun_out/last_suite.log; tail -3 gpurun_out/last_suite.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/last_smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/last_smoke.log
sed -i 's/r[0-9]_launches/last_launches/g; s/r[0-9]_prof/last_prof/g; s/final_launches/last_launches/g; s/final_prof/last_prof/g' tools/r2_profile.sh; bash tools/r2_profile.sh
python tools/ncu_kernels_json.py gpurun_out/last_prof.ncu-rep > gpurun_out/last_ncu_kernels.json 2> gpurun_out/last_ncu_kernels.err && cp gpurun_out/last_ncu_kernels.json profiles/r2_ncu_kernels.json
head -c 600 gpurun_out/last_ncu_kernels.json; echo
timeout 900 python bench.py > gpurun_out/last_bench.json 2> gpurun_out/last_bench.err; echo "bench rc=$?"; python - <<'PY'
import json
d = json.loads(open("gpurun_out/last_bench.json").read().strip().splitlines()[-1])
print("ms/step", d["ms_per_step"], "value %.4g" % d["value"], "traffic", d["roofline"]["traffic"], d["roofline"].get("traffic_source"))
print("issue", d["roofline"].get("issue"))
print("e2e", d["e2e"]["ms_per_step"])
PY
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/last_bench_reference.json 2> gpurun_out/last_bench_reference.err; tail -c 300 gpurun_out/last_bench_reference.json
