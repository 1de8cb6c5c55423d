#!/bin/bash
# Round-end evidence run on one B200 (under gpurun): GPU tests, smoke, benches, reference arm, ncu launch list of the
# bench command and one ncu --set full capture of representative GEMM launches. Outputs under gpurun_out/final_*.
set -u
O=gpurun_out
timeout 600 python -m pytest tests -q -m gpu > $O/final_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 $O/final_pytest_gpu.log
GP_EXTENDED_TESTS=1 timeout 300 python -m pytest tests/test_gpu_trace.py -q -s > $O/final_pytest_gpu_extended.log 2>&1; echo "extended rc=$?"; grep -i "engine mixed\|passed\|failed" $O/final_pytest_gpu_extended.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/final_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/final_smoke.log
timeout 300 python bench.py --steps 20 --warmup 5 > $O/final_bench_n1_bf16x3.json 2> $O/final_bench_n1_bf16x3.err; echo "bench x3 rc=$?"
GP_PRECISION=bf16 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/final_bench_n1_bf16.json 2> $O/final_bench_n1_bf16.err; echo "bench bf16 rc=$?"
GP_FAKE_PRECISION=bf16x3 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/final_bench_n1_bf16x3_all_fake_passes.json 2> /dev/null; echo "bench x3-everywhere rc=$?"
timeout 200 python tools/bench_loops.py --steps 20 --warmup 5 > $O/final_bench_loops.jsonl 2> /dev/null; echo "bench loops rc=$?"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/final_bench_reference.json 2> $O/final_bench_reference.err; echo "reference arm rc=$?"
for p in bf16x3 bf16; do timeout 120 python tools/step_breakdown.py --batch 1024 --precision $p --gemms --out $O/final_breakdown_${p}.log > /dev/null 2>&1; done
timeout 120 python tools/prof_gemm.py --reps 10 > $O/final_gemm_microbench.log 2>&1
# ncu: launch list of the bench command (eager launches so every kernel is its own node), then one full capture
timeout 200 python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline > $O/final_plain_for_ncu.log 2>&1 && \
timeout 500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 1500 -c 1000 \
    --csv --log-file $O/final_ncu_launches.csv python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline > $O/final_ncu_launches.log 2>&1
echo "ncu launches rc=$?"
timeout 100 python tools/prof_gemm.py --reps 1 --cases d1_fwd_stats,d1_fwd_x3,d2_wgrad,img_fwd > $O/final_plain_for_ncu_full.log 2>&1 && \
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_gemm -c 8 -o $O/final_prof_gemm \
    python tools/prof_gemm.py --reps 1 --cases d1_fwd_stats,d1_fwd_x3,d2_wgrad,img_fwd > $O/final_ncu_full.log 2>&1
echo "ncu full rc=$?"
