#!/bin/bash
# Round-2 evidence run on one B200 (under gpurun): GPU tests, smoke, benches of every configuration, reference arm, ncu
# launch list + DRAM traffic of the headline bench command (the BENCHED precision policy) and one ncu --set full capture of
# representative GEMM launches, compute-sanitizer logs. Outputs under gpurun_out/r02_*; copy what is to be judged into
# profiles/. Usage: bash tools/final_artifacts.sh [tests|bench|ncu|sanitize ...]   (default: everything)
set -u
O=gpurun_out
mkdir -p $O
WHAT="${*:-tests bench ncu sanitize}"
has() { [[ " $WHAT " == *" $1 "* ]]; }

if has tests; then
  rm -f $O/parity_table.jsonl
  timeout 1500 python -m pytest tests -q -m gpu -rs -s --durations=15 > $O/r02_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r02_pytest_gpu.log
  timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/r02_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/r02_smoke.log
fi
if has bench; then
  for c in cfg2 cfg3 cfg4 cfg5; do
    timeout 400 python bench.py --config $c --steps 20 --warmup 5 > $O/r02_bench_n1_$c.json 2> $O/r02_bench_n1_$c.err; echo "bench $c rc=$?"
  done
  GP_PRECISION=bf16 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/r02_bench_n1_cfg2_bf16.json 2> /dev/null; echo "bench bf16 rc=$?"
  timeout 300 python bench.py --global-batch 128 --steps 50 --warmup 5 --no-cpu-baseline > $O/r02_bench_n1_cfg2_batch128.json 2> /dev/null; echo "bench b128 rc=$?"
  timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > $O/r02_bench_reference_cfg2.json 2> $O/r02_bench_reference.err; echo "reference arm rc=$?"
  timeout 200 python tools/step_breakdown.py --batch 1024 --precision bf16x3 --gemms --out $O/r02_breakdown_b1024.log > /dev/null 2>&1; echo "breakdown rc=$?"
  timeout 200 python tools/step_breakdown.py --batch 128 --precision bf16x3 --gemms --out $O/r02_breakdown_b128.log > /dev/null 2>&1
  timeout 120 python tools/prof_gemm.py --reps 10 > $O/r02_gemm_microbench.log 2>&1
fi
if has ncu; then
  # launch list of the bench command (eager launches so every kernel is its own node) with DRAM bytes per launch
  timeout 200 python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline > $O/r02_plain_for_ncu.json 2> /dev/null && \
  timeout 700 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 1200 -c 1400 \
      --csv --log-file $O/r02_ncu_launches_n1_cfg2.csv python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline > $O/r02_ncu_launches.log 2>&1
  echo "ncu launches rc=$?"
  LABEL=$(python -c "import json;print(json.load(open('$O/r02_plain_for_ncu.json'))['config']['precision'])")
  python tools/ncu_traffic.py $O/r02_ncu_launches_n1_cfg2.csv cfg2 "$LABEL" 1024 profiles/r02_ncu_launches_n1_cfg2.csv > $O/r02_ncu_traffic.log 2>&1; echo "traffic rc=$?"
  python tools/ncu_summary.py $O/r02_ncu_launches_n1_cfg2.csv > $O/r02_ncu_launches_n1_cfg2_summary.txt 2>&1
  timeout 100 python tools/prof_gemm.py --reps 1 --cases d1_fwd_stats,d1_fwd_x3,d2_wgrad,img_fwd > $O/r02_plain_for_ncu_full.log 2>&1 && \
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:conv_gemm -c 8 -o $O/r02_prof_gemm \
      python tools/prof_gemm.py --reps 1 --cases d1_fwd_stats,d1_fwd_x3,d2_wgrad,img_fwd > $O/r02_ncu_full.log 2>&1
  echo "ncu full rc=$?"
fi
if has sanitize; then
  timeout 200 python tools/sanitize_cases.py --quick > $O/r02_sanitize_plain.log 2>&1; echo "sanitize cases (no tool) rc=$?"
  # compute-sanitizer itself is closed on this pool (profiles/r02_compute_sanitizer_closed_on_pool.log)
  timeout 200 python tools/check_pair.py > $O/r02_check_pair.log 2>&1; echo "check_pair rc=$?"
  timeout 200 python tools/check_wgrad_pair.py > $O/r02_check_wgrad_pair.log 2>&1; echo "check_wgrad_pair rc=$?"
  timeout 200 python tools/check_wide_flat_wgrad.py > $O/r02_check_wide_flat_wgrad.log 2>&1; echo "check_wide_flat rc=$?"
fi
