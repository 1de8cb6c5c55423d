#!/bin/bash
# G(z1) next to the real-image pass (config.g_ahead): engine / loop / trace tests, then batch-128 / 256 benches on / off
set -u
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_engine.py tests/test_gpu_weight_stage.py tests/test_gpu_zz_loops.py tests/test_gpu_trace.py tests/test_gpu_parity_bars.py -q -m gpu > $O/r02_ga_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r02_ga_pytest_gpu.log
for w in 1 0; do
  GP_G_AHEAD=$w timeout 300 python bench.py --global-batch 128 --steps 100 --warmup 10 --no-cpu-baseline > $O/r02_ga${w}_cfg2_b128.json 2> $O/r02_ga${w}_b128.err; echo "b128 ga=$w rc=$?"
  GP_G_AHEAD=$w timeout 300 python bench.py --global-batch 256 --steps 100 --warmup 10 --no-cpu-baseline > $O/r02_ga${w}_cfg2_b256.json 2> /dev/null; echo "b256 ga=$w rc=$?"
  GP_G_AHEAD=$w GP_WGRAD_STREAM=$w timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > $O/r02_ga${w}_cfg2.json 2> /dev/null; echo "cfg2 ga=ws=$w rc=$?"
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r02_ga?_*.json')):
    try:
        d = json.load(open(f)); print(f, round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['ms_per_step'], 3))
    except Exception as e:
        print(f, 'unreadable', e)
PY
