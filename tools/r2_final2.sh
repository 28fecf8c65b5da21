#!/bin/bash
# closing pass on one B200 after the last kernel edit: (suite unless SKIP_SUITE=1) + smoke, ncu launch list + full capture of the
# step kernels, the capture's per-launch numbers written where bench.py looks for them, then the default bench line.
cd "$(dirname "$0")/.." && mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
if [ -z "$SKIP_SUITE" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/final2_suite.log 2>&1; echo "suite rc=$?" | tee -a gpurun_out/final2_suite.log; tail -3 gpurun_out/final2_suite.log
fi
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final2_smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/final2_smoke.log
BENCH="python bench.py --no-e2e --no-cpu-baseline --no-rows"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 160 --csv --log-file gpurun_out/final2_launches.csv $BENCH --steps 16 --warmup 4 > gpurun_out/final2_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"render_|preprocess_|emit_|tile_" -s 16 -c 16 -f -o gpurun_out/final2_prof $BENCH --steps 3 --warmup 3 > gpurun_out/final2_prof.log 2>&1
ls -la gpurun_out/final2_prof.ncu-rep gpurun_out/final2_launches.csv; tail -2 gpurun_out/final2_prof.log
python tools/ncu_kernels_json.py gpurun_out/final2_prof.ncu-rep > gpurun_out/final2_ncu_kernels.json 2> gpurun_out/final2_ncu_kernels.err && cp gpurun_out/final2_ncu_kernels.json profiles/r2_ncu_kernels.json
head -c 400 gpurun_out/final2_ncu_kernels.json; echo
timeout 900 python bench.py > gpurun_out/final2_bench.json 2> gpurun_out/final2_bench.err; echo "bench rc=$?"; python - <<'PY'
import json
d = json.loads(open("gpurun_out/final2_bench.json").read().strip().splitlines()[-1])
print("ms/step", d["ms_per_step"], "value %.4g" % d["value"], "traffic", d["roofline"]["traffic"], d["roofline"].get("traffic_source"))
print("kernels", d["roofline"].get("kernels_ms_per_view"))
print("issue", d["roofline"].get("issue"))
print("e2e", d["e2e"]["ms_per_step"])
print("other", {k: v.get("ms_per_step") for k, v in d.get("other_shapes", {}).items()})
PY
