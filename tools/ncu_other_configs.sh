#!/bin/bash
# ncu launch lists (duration + DRAM bytes per launch) of the cfg3 / cfg4 / cfg5 bench commands, so that their
# `roofline.traffic` is measured like cfg2's (tools/final_artifacts.sh). Each capture follows a plain run of the same
# command that exited 0. Post-process here or on the CPU box:
#   python tools/ncu_traffic.py gpurun_out/r02_ncu_launches_n1_cfgN.csv cfgN "<precision label>" <batch> profiles/r02_ncu_launches_n1_cfgN_tail.csv
set -u
O=gpurun_out
mkdir -p $O
for c in ${*:-cfg3 cfg4 cfg5}; do
  timeout 200 python bench.py --config $c --steps 1 --warmup 3 --no-graph --no-cpu-baseline > $O/r02_plain_for_ncu_$c.json 2> /dev/null && \
  timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
      --csv --log-file $O/r02_ncu_launches_n1_$c.csv python bench.py --config $c --steps 1 --warmup 3 --no-graph --no-cpu-baseline > $O/r02_ncu_launches_$c.log 2>&1
  echo "ncu launches $c rc=$?"; ls -la $O/r02_ncu_launches_n1_$c.csv
done
