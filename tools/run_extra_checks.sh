set -x
for cfg in "500000 512 init" "500000 512 trained" "2000000 1600 trained" "100000 800 init"; do
  set -- $cfg
  python bench.py --points $1 --res $2 --opacity $3 --steps 100 --warmup 5 --no-e2e --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('CFG', d['config']['workload'], '| ms/step', round(d['ms_per_step'],4), '| splats/s %.3e' % d['value'], '| n_inst', d['config']['n_inst_view0'], '| step frac', round(d['roofline']['step']['frac'],4), '| kernels', d['roofline']['kernels_ms'])"
done
SAN="compute-sanitizer --tool memcheck --error-exitcode 9 --launch-timeout 0"
timeout 500 $SAN python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "raw_deg1 or raw_deg0 or padded_sh" 2>&1 | tail -3
timeout 400 $SAN python -m pytest tests/test_knn_gpu.py -m gpu -q -x -k "small_sets or line or duplicates" 2>&1 | tail -3
timeout 400 $SAN python -m pytest tests/test_fields_gpu.py -m gpu -q -x -k "reference_outputs or model_style" 2>&1 | tail -3
timeout 300 $SAN python -m pytest tests/test_stage1_gpu.py -m gpu -q -x -k "adam" 2>&1 | tail -3
