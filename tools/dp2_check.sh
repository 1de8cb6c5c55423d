#!/bin/bash
# 2-GPU run of the step driver (NCCL ZeRO-1 exchange, peer-memory SyncBN, CUDA graph) with the weight-gradient side
# stream on / off at the 8-GPU shard size (128 per GPU), then the default strong-scaling line at global 1024
set -u
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
for w in 1 0; do
  GP_WGRAD_STREAM=$w GP_G_AHEAD=$w timeout 300 $TR --master-port 2951$w bench.py --gpus 2 --global-batch 256 --steps 100 --warmup 10 --no-cpu-baseline > $O/r02_dp2_ws${w}_b256.json 2> $O/r02_dp2_ws${w}_b256.err; echo "N=2 global 256 ws=$w rc=$?"
done
timeout 300 $TR --master-port 29513 bench.py --gpus 2 --steps 50 --warmup 5 --no-cpu-baseline > $O/r02_dp2_cfg2.json 2> $O/r02_dp2_cfg2.err; echo "N=2 cfg2 rc=$?"
timeout 300 $TR --master-port 29514 tools/dp_check.py > $O/r02_dp_check_n2.log 2>&1; echo "dp_check rc=$?"; tail -4 $O/r02_dp_check_n2.log
timeout 300 $TR --master-port 29515 tools/dp_engine_check.py > $O/r02_dp_engine_check_n2.log 2>&1; echo "dp_engine_check rc=$?"; tail -5 $O/r02_dp_engine_check_n2.log
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r02_dp2_*.json')):
    try:
        d = json.load(open(f)); print(f, round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['ms_per_step'], 3), d['e2e'].get('last_logged'))
    except Exception as e:
        print(f, 'unreadable', e)
PY
