#!/bin/bash
# coalesced weight-gradient reductions + whole-tile schedule (conv_gemm): correctness cases, then benches with GP_WGRAD_WHOLE on / off
set -u
O=gpurun_out
mkdir -p $O
timeout 300 python tools/sanitize_cases.py --quick > $O/r02_wg_sanitize.log 2>&1; echo "sanitize cases rc=$?"; tail -2 $O/r02_wg_sanitize.log
timeout 200 python tools/check_wide_flat_wgrad.py > $O/r02_wg_wide_flat.log 2>&1; echo "wide flat rc=$?"
timeout 900 python -m pytest tests -q -m gpu -x > $O/r02_wg_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 $O/r02_wg_pytest_gpu.log
for w in 1 0; do
  GP_WGRAD_WHOLE=$w timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > $O/r02_wg${w}_cfg2.json 2> $O/r02_wg${w}_cfg2.err; echo "cfg2 whole=$w rc=$?"
  GP_WGRAD_WHOLE=$w timeout 300 python bench.py --global-batch 128 --steps 100 --warmup 10 --no-cpu-baseline > $O/r02_wg${w}_cfg2_b128.json 2> /dev/null; echo "b128 whole=$w rc=$?"
done
GP_WGRAD_WHOLE=1 timeout 200 python tools/step_breakdown.py --batch 128 --precision bf16x3 --gemms --out $O/r02_wg1_breakdown_b128.log > /dev/null 2>&1
GP_WGRAD_WHOLE=1 timeout 200 python tools/step_breakdown.py --batch 1024 --precision bf16x3 --gemms --out $O/r02_wg1_breakdown_b1024.log > /dev/null 2>&1
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r02_wg?_*.json')):
    try:
        d = json.load(open(f)); print(f, round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['ms_per_step'], 3), 'gemm', round(d['roofline']['gemm_ms_per_step'], 3))
    except Exception as e:
        print(f, 'unreadable', e)
PY
grep -n "conv_wgrad" $O/r02_wg1_breakdown_b128.log | tail -12
