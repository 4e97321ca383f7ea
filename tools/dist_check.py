"""Multi-GPU check of the sharded path: torchrun --nproc-per-node N tools/dist_check.py n [p] [reps]
Compares the N-GPU -2logL with the single-GPU value computed on rank 0 (if it fits)."""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from varycoef_b200 import SVCObjective  # noqa: E402
from varycoef_b200.dist import distributed_objective, process_grid  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 6000
p = int(sys.argv[2]) if len(sys.argv) > 2 else 3
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
single = (sys.argv[4] != "0") if len(sys.argv) > 4 else True
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
rng = np.random.default_rng(1)
L = 10.0 * np.sqrt(n / 1000.0)
locs = rng.uniform(0, L, (n, 2))
X = np.column_stack([np.ones(n), rng.standard_normal((n, p - 1))])
y = X @ np.arange(1, p + 1) + rng.standard_normal(n) * 1.8
theta = np.array([1.3, 0.7, 2.1, 0.4, 0.9, 1.1, 1.7, 0.3, 2.5, 0.2][:2 * p] + [0.35])
obj = distributed_objective(y, X, locs, profileLik=True, device=local)
for r in range(reps):
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.time()
    res = obj.n2LL(theta, parts=True)
    torch.cuda.synchronize()
    dt = time.time() - t0
    tm = obj.timings()
    if rank == 0:
        print("dist %dx%d n=%d p=%d n2ll=%.12g mu=%s wall=%.3fs build=%.1fms chol=%.1fms fin=%.2fms  chol %.2f TFLOP/s aggregate" % (
            *process_grid(world), n, p, res["n2ll"], np.array2string(res["mu"], precision=10), dt, tm["build_ms"],
            tm["chol_ms"], tm["finalize_ms"], n ** 3 / 3 / (tm["chol_ms"] * 1e-3) / 1e12), flush=True)
# fixed-mu mode
v2 = obj.n2LL(np.concatenate([theta, res["mu"] * 1.01]), profile=False)
obj.close()
if rank == 0 and single:
    ref = SVCObjective(y, X, locs, profileLik=True, device=local)
    r1 = ref.n2LL(theta, parts=True)
    w2 = ref.n2LL(np.concatenate([theta, res["mu"] * 1.01]), profile=False)
    ref.close()
    print("single n2ll=%.12g  rel diff=%.3e  mu rel diff=%.3e  fixed-mu rel diff=%.3e" % (
        r1["n2ll"], abs(r1["n2ll"] - res["n2ll"]) / abs(r1["n2ll"]),
        np.max(np.abs(r1["mu"] - res["mu"]) / np.abs(r1["mu"])), abs(w2 - v2) / abs(w2)), flush=True)
    bad = max(abs(r1["n2ll"] - res["n2ll"]) / abs(r1["n2ll"]), abs(w2 - v2) / abs(w2)) > 1e-13
    print("DIST_CHECK " + ("FAIL" if bad else "OK"), flush=True)
dist.barrier()
dist.destroy_process_group()
