#!/bin/bash
# 8-GPU measurements of the sharded path (C4) and of the default bench line; outputs under gpurun_out/
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 8 "${@:2}"; }
for R in 16 24 10; do
  VC_DIST_RESERVE=$R run 2952$((R % 10)) --workload c4 --steps 2 --warmup 1 > gpurun_out/r2_c4_n8_R$R.json 2> gpurun_out/r2_c4_n8_R$R.err
  python -c "
import json; d=json.load(open('gpurun_out/r2_c4_n8_R$R.json')); print('R=$R', d['ms_per_step'], d['roofline']['achieved'], d['roofline']['frac'], d['n2ll_sample'])"
done
run 29540 --steps 3 --warmup 2 > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_n8.json')); print('bench n8', d['value'], d['ms_per_step']); print(json.dumps(d.get('sharded'))[:1500])"
python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -3
