#!/bin/bash
# 8-GPU measurements of the sharded path (C4); outputs under gpurun_out/
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 8 "${@:2}"; }
i=0
for cfg in "VC_DIST_RESERVE=8" "VC_DIST_RESERVE=8 VC_DIST_SAFETY=100"; do
  i=$((i+1))
  env $cfg bash -c "$(declare -f run); run 2952$i --workload c4 --steps 2 --warmup 1" > gpurun_out/r2_c4_n8_x$i.json 2> gpurun_out/r2_c4_n8_x$i.err
  python -c "
import json; d=json.load(open('gpurun_out/r2_c4_n8_x$i.json')); print('$cfg', d['ms_per_step'], d['roofline']['achieved'], d['roofline']['frac'], d['n2ll_sample'])"
done
