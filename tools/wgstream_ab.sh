#!/bin/bash
# weight-gradient GEMMs on a side stream (config.wgrad_side): full GPU suite (auto = on at the tests' small batches), then
# benches with GP_WGRAD_STREAM=1 / 0 on the same box
set -u
O=gpurun_out
mkdir -p $O
rm -f $O/parity_table.jsonl
timeout 900 python -m pytest tests -q -m gpu > $O/r02_ws_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r02_ws_pytest_gpu.log
for w in 1 0; do
  GP_WGRAD_STREAM=$w timeout 300 python bench.py --global-batch 128 --steps 100 --warmup 10 --no-cpu-baseline > $O/r02_ws${w}_cfg2_b128.json 2> $O/r02_ws${w}_b128.err; echo "b128 ws=$w rc=$?"
  GP_WGRAD_STREAM=$w timeout 300 python bench.py --global-batch 256 --steps 100 --warmup 10 --no-cpu-baseline > $O/r02_ws${w}_cfg2_b256.json 2> /dev/null; echo "b256 ws=$w rc=$?"
  GP_WGRAD_STREAM=$w timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > $O/r02_ws${w}_cfg2.json 2> /dev/null; echo "cfg2 ws=$w rc=$?"
  GP_WGRAD_STREAM=$w timeout 300 python bench.py --config cfg5 --steps 30 --warmup 5 --no-cpu-baseline > $O/r02_ws${w}_cfg5.json 2> /dev/null; echo "cfg5 ws=$w rc=$?"
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r02_ws?_*.json')):
    try:
        d = json.load(open(f)); print(f, round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['ms_per_step'], 3))
    except Exception as e:
        print(f, 'unreadable', e)
PY
