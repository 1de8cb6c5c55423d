#!/bin/bash
# A/B of programmatic dependent launch (csrc/common.h, GP_PDL): GPU tests with it on, then the same benches with it on / off.
set -u
O=gpurun_out
mkdir -p $O
rm -f $O/parity_table.jsonl
GP_PDL=1 timeout 900 python -m pytest tests -q -m gpu -x > $O/r02_pdl_pytest_gpu.log 2>&1; echo "pytest (GP_PDL=1) rc=$?"; tail -2 $O/r02_pdl_pytest_gpu.log
for pdl in 1 0; do
  GP_PDL=$pdl timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > $O/r02_pdl${pdl}_cfg2.json 2> $O/r02_pdl${pdl}_cfg2.err; echo "cfg2 pdl=$pdl rc=$?"
  GP_PDL=$pdl timeout 300 python bench.py --global-batch 128 --steps 100 --warmup 10 --no-cpu-baseline > $O/r02_pdl${pdl}_cfg2_b128.json 2> /dev/null; echo "b128 pdl=$pdl rc=$?"
  GP_PDL=$pdl timeout 300 python bench.py --config cfg3 --steps 50 --warmup 5 --no-cpu-baseline > $O/r02_pdl${pdl}_cfg3.json 2> /dev/null; echo "cfg3 pdl=$pdl rc=$?"
  GP_PDL=$pdl timeout 300 python bench.py --config cfg4 --steps 20 --warmup 5 --no-cpu-baseline > $O/r02_pdl${pdl}_cfg4.json 2> /dev/null; echo "cfg4 pdl=$pdl rc=$?"
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r02_pdl?_*.json')):
    try:
        d = json.load(open(f)); print(f, round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['ms_per_step'], 3))
    except Exception as e:
        print(f, 'unreadable', e)
PY
