#!/usr/bin/env python
"""bench.py -- headline benchmark of the -2 logL objective (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c3]

One "step" = one evaluation of -2 logL (fused covariance build -> bordered DMMA Cholesky -> fused
GLS / quadratic form / log-det) on synthetic sample_SVCdata-style inputs.  Default workload is the
headline configuration C3 of BASELINE.json: n = 50 000, p = q = 3, exponential covariance, fp64.

N > 1 (launched by torch.distributed.run, one rank per GPU): the objective shards across the
finite-difference stencil optim evaluates -- every rank holds a replica of the O(n) inputs and
evaluates its own theta; there is no collective on the data path ("scaling": "weak").
The within-evaluation sharded path (n = 150k, 2D block-cyclic + NCCL) is `--workload c4`.

`--impl reference` times the reference's CPU op sequence (numpy/LAPACK restatement in oracle/,
the reference itself is R and cannot run here) on the host cores, on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# NCCL prints its version banner on stdout at the default debug level of this image; the contract
# is ONE JSON line on stdout, so keep NCCL to warnings unless the caller asked for more
if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
    os.environ["NCCL_DEBUG"] = "WARN"
# ... and route everything any library writes to fd 1 to stderr; the JSON line goes to the real stdout
_REAL_STDOUT = os.fdopen(os.dup(1), "w")
os.dup2(2, 1)


def emit(line: dict):
    _REAL_STDOUT.write(json.dumps(line) + "\n")
    _REAL_STDOUT.flush()

WORKLOADS = {
    # name: (n, p=q, cov, taper, seed offset, description)
    "c1": (1000, 2, "exp", None, 1, "C1 n=1000 p=q=2 exp"),
    "c2": (10000, 3, "mat32", None, 2, "C2 n=10000 p=q=3 matern-3/2"),
    "c3": (50000, 3, "exp", None, 3, "C3 n=50000 p=q=3 exp (headline)"),
    "c5": (20000, 3, "exp", 0.6, 5, "C5 n=20000 p=q=3 exp + Wendland-1 taper delta=0.6"),
    "c4": (150000, 5, "exp", None, 4, "C4 n=150000 p=q=5 exp, 2D block-cyclic over N GPUs"),
}
TRUE_VAR = [2.0, 1.0, 1.5, 1.0, 0.5]
TRUE_SCALE = [3.0, 1.0, 2.0, 1.5, 2.5]
TRUE_MEAN = [1.0, 2.0, 0.5, -1.0, 0.0]
# DRAM bytes per evaluation from ncu (dram__bytes_read.sum + dram__bytes_write.sum), profiles/r01_traffic_n50k.md
NCU_TRAFFIC_BYTES = {"c3": 6.163e11}           # potrf + trsm + syrk kernels of one C3 evaluation
NCU_BUILD_TRAFFIC_BYTES = {"c3": 1.002e10}     # k_build_fast + k_border (algorithmic: 1.0e10 written)
METRIC = "-2logL evals/sec (fp64) at n=50k q=3"
UNIT = "evals/s"


def make_inputs(n, p, seed_off):
    """sample_SVCdata-style synthetic inputs: locs ~ U(0, L)^2 at constant
    density 10 / unit^2, X = [1, N(0,1)], W = X; y = X mu + N(0, .) (a single evaluation does not
    depend on y being a GP draw)."""
    rng = np.random.default_rng(20261019 + seed_off)
    L = 10.0 * np.sqrt(n / 1000.0)
    locs = rng.uniform(0.0, L, size=(n, 2))
    X = np.column_stack([np.ones(n), rng.standard_normal((n, p - 1))])
    y = X @ np.array(TRUE_MEAN[:p]) + rng.standard_normal(n) * 1.8
    return np.asfortranarray(locs), np.asfortranarray(X), y, L


def theta_stencil(n, p, q, L, y):
    """The 1 + 2 (2q+1) points optim's central-difference gradient evaluates around the default
    init (R-pkg/R/utils.R:376-381), step 1e-3 * parscale, plus theta_true.  med_dist of a uniform
    square is 0.5120 L (exact median needs the O(n^2) selection pass; not part of the hot path)."""
    med = 0.5120 * L
    yv = float(np.var(y, ddof=1))
    init = np.array([med / 4.0, yv / (q + 1)] * q + [yv / (q + 1)])
    pts = [init.copy()]
    for i in range(init.size):
        for s in (+1.0, -1.0):
            t = init.copy()
            t[i] += s * 1e-3 * abs(init[i])
            pts.append(t)
    true = []
    for k in range(q):
        true += [TRUE_SCALE[k], TRUE_VAR[k]]
    pts.append(np.array(true + [0.25]))
    return np.array(pts)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); pw.append(float(r[3]))
                for nm, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        busy = [s for s, w in zip(sm, pw) if w > 0.5 * max(pw)] or sm
        return {"sm_mhz": float(np.median(busy)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(pw)),
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference's op sequence (oracle port), bounded sample, extrapolated to the workload
# ------------------------------------------------------------------------------------------------
def cpu_eval_timed(n_s, p, cov, seed_off, faithful=True):
    """One evaluation of the reference's dense n2LL op sequence at n_s on the host BLAS threads.
    Returns (t_build, t_cubic): Sigma_y build (O(n^2)) and chol + solves + determinant (O(n^3))."""
    from oracle import svc_oracle as o
    locs, X, y, L = make_inputs(n_s, p, seed_off)
    th = theta_stencil(n_s, p, p, L, y)[0]
    t0 = time.perf_counter()
    D = o.own_dist(locs)
    t_dist = time.perf_counter() - t0           # once per fit in the reference: not charged
    t0 = time.perf_counter()
    S = o.sigma_y(th, D, X, cov)
    t_build = time.perf_counter() - t0
    del S
    t0 = time.perf_counter()
    v = o.n2ll(th, D, X, X, y, cov, profile=True, faithful=faithful)
    t_all = time.perf_counter() - t0
    return t_build, max(t_all - t_build, 1e-9), t_dist, v


def cpu_arm(n, p, cov, seed_off, n_s, reps):
    from threadpoolctl import threadpool_info
    tb, tc = [], []
    for _ in range(reps):
        b, c, _, _ = cpu_eval_timed(n_s, p, cov, seed_off, faithful=True)
        tb.append(b); tc.append(c)
    tb, tc = float(np.median(tb)), float(np.median(tc))
    t_full = tb * (n / n_s) ** 2 + tc * (n / n_s) ** 3
    threads = max([i.get("num_threads", 1) for i in threadpool_info()] + [1])
    return {"value": 1.0 / t_full, "unit": UNIT, "cores": int(threads), "kind": "port",
            "sample": ("faithful R op sequence (Sigma_y loop, dpotrf, 3x dgesv on the triangular factor, log-det) "
                       "at n_s=%d, p=q=%d, %s: build %.3fs (x(n/n_s)^2) + cubic part %.3fs (x(n/n_s)^3) "
                       "extrapolated to n=%d; host has %d logical cores"
                       % (n_s, p, cov, tb, tc, n, os.cpu_count())),
            "sample_seconds": tb + tc}


def run_reference(args):
    n, p, cov, taper, seed_off, desc = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_s = args.cpu_sample_n
    for _ in range(args.warmup if args.warmup < 2 else 1):
        cpu_eval_timed(min(n_s, 2000), p, cov, seed_off)
    t0 = time.perf_counter()
    arm = cpu_arm(n, p, cov, seed_off, n_s, max(1, args.steps))
    wall = time.perf_counter() - t0
    line = {"impl": "reference", "metric": METRIC, "value": arm["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 / arm["value"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": {"workload": desc, "n": n, "p": p, "q": p, "cov": cov},
            "cpu_baseline": arm,
            "e2e": {"value": arm["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "wall_s": wall}
    emit(line)


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import ctypes as C
    from varycoef_b200 import SVCObjective, _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the product has no CPU path)")
    torch.cuda.set_device(local)
    n, p, cov, taper, seed_off, desc = WORKLOADS[args.workload]
    if args.n:
        n = args.n
    q = p
    locs, X, y, L = make_inputs(n, p, seed_off)
    thetas = theta_stencil(n, p, q, L, y)
    K, Wm = args.steps, args.warmup

    # pinned host copies of the O(n) inputs for the end-to-end leg
    def pinned(a):
        t = torch.empty(a.shape[::-1] if a.ndim == 2 else a.shape, dtype=torch.float64, pin_memory=True)
        v = t.numpy()
        if a.ndim == 2:
            v = v.T            # column-major view of the pinned buffer
        v[...] = a
        return t, v
    _tl, locs_p = pinned(locs)
    _tx, X_p = pinned(X)
    _ty, y_p = pinned(y)

    obj = SVCObjective(y, X, locs, cov_name=cov, tapering=taper, profileLik=True, device=local)
    lib = _lib.lib()
    peak_tf = C.c_double(0.0)
    lib.vc_probe_dmma_peak(local, 20000, C.byref(peak_tf))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def theta_of(step):
        return thetas[(step * world + rank) % len(thetas)]

    # ---- device-resident leg ("value") ----
    for s in range(Wm):
        obj.n2LL(theta_of(s))
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    l0 = obj.launch_count
    t0 = time.perf_counter()
    phase = {"build_ms": 0.0, "chol_ms": 0.0, "finalize_ms": 0.0}
    vals = []
    for s in range(K):
        vals.append(obj.n2LL(theta_of(Wm + s)))
        tm = obj.timings()
        for k in phase:
            phase[k] += tm[k]
    barrier()
    elapsed = time.perf_counter() - t0
    launches = obj.launch_count - l0
    clocks = sampler.stop() if rank == 0 else None

    # ---- end-to-end leg: host buffers in, scalar out, copies inside the timed region ----
    barrier()
    t0 = time.perf_counter()
    for s in range(K):
        obj.set_data(y=y_p, X=X_p, locs=locs_p, W=X_p)
        obj.n2LL(theta_of(Wm + s))
    barrier()
    elapsed_e2e = time.perf_counter() - t0

    if dist is not None:
        t = torch.tensor([elapsed, elapsed_e2e], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed, elapsed_e2e = float(t[0]), float(t[1])
        ph = torch.tensor([phase["build_ms"], phase["chol_ms"], phase["finalize_ms"], float(launches)],
                          dtype=torch.float64, device="cuda")
        dist.all_reduce(ph, op=dist.ReduceOp.SUM)
        launches_total = int(ph[3].item())
    else:
        launches_total = launches
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    n_pad = obj.padded_n
    chol_s = phase["chol_ms"] / K * 1e-3
    build_s = phase["build_ms"] / K * 1e-3
    chol_flops = n ** 3 / 3.0
    build_bytes = 4.0 * n * (n + 1)            # lower triangle is what the build writes
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    hbm_src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "6650 GB/s (of fallback)"
    roof = {"bound": "tensor", "kernel": "k_syrk_ws (persistent TMA-fed DMMA trailing SYRK) inside the bordered Cholesky",
            "achieved": chol_flops / chol_s / 1e12, "peak": peak_tf.value, "unit": "TFLOP/s",
            "frac": chol_flops / chol_s / 1e12 / peak_tf.value if peak_tf.value > 0 else None,
            "traffic": NCU_TRAFFIC_BYTES.get(args.workload) if n == WORKLOADS[args.workload][0] else None,
            "traffic_note": "DRAM read + write bytes of all Cholesky kernels of ONE evaluation (per evaluation, like "
                            "`achieved`), ncu dram__bytes_{read,write}.sum, profiles/r01_traffic_n50k.md: 616 GB in "
                            "1.21 s = 8 % of DRAM peak -- the factorisation is DMMA-bound (pipe 95.7 % active in "
                            "k_syrk_ws, profiles/r01_syrk_ws_n20k.md); the C trapezoid alone accounts for ~400 GB",
            "peak_source": "fp64 DMMA m8n8k4 register-resident probe measured live on this GPU (vc_probe_dmma_peak); "
                           "MEASURED_PEAKS.json has no fp64 entry",
            "algorithmic": "n^3/3 flops per evaluation / CUDA-event time of the whole factorisation (panels included)"}
    roof_build = {"bound": "hbm", "kernel": "k_build_fast (fused distance + covariance)", "achieved": build_bytes / build_s / 1e9,
                  "peak": hbm_peak, "unit": "GB/s", "frac": build_bytes / build_s / 1e9 / hbm_peak,
                  "traffic": NCU_BUILD_TRAFFIC_BYTES.get(args.workload) if n == WORKLOADS[args.workload][0] else None, "peak_source": hbm_src,
                  "algorithmic": "4 n (n+1) bytes written (lower triangle) / CUDA-event time of build + border"}
    h2d = 8 * n * (2 + p + q + 1) + 8 * (2 * q + 1)
    d2h = 8 * 132
    line = {
        "metric": METRIC, "value": K * world / elapsed, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": Wm,
        "ms_per_step": elapsed / K * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": desc, "n": n, "n_padded": n_pad, "p": p, "q": q, "cov": cov, "tapering": taper,
                   "mu_mode": "profile/GLS", "theta": "optim central-difference stencil around default init, cycled",
                   "parallelism": "theta-parallel replicas, no collective" if world > 1 else "single GPU",
                   "l2": "inputs larger than L2: every evaluation rewrites and refactors the %.1f GB matrix" % (
                       8.0 * n_pad * (n_pad + 128) / 1e9)},
        "clocks": clocks,
        "e2e": {"value": K * world / elapsed_e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": launches_total,
        "roofline": roof, "roofline_build": roof_build,
        "phases_ms": {k: v / K for k, v in phase.items()},
        "n2ll_sample": vals[:2],
    }
    if world == 1 and not args.no_cpu:
        line["cpu_baseline"] = cpu_arm(n, p, cov, seed_off, args.cpu_sample_n, 1)
    emit(line)
    obj.close()
    if dist is not None:
        dist.destroy_process_group()


def run_dist(args):
    """--workload c4: ONE evaluation sharded 2D block-cyclic over all ranks (n = 150 000 does not fit
    one GPU).  value = whole-job evaluations per second; strong scaling in N."""
    import torch
    import torch.distributed as dist
    from varycoef_b200.dist import distributed_objective, process_grid
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29533")
        dist.init_process_group("gloo", rank=0, world_size=1)
    n, p, cov, taper, seed_off, desc = WORKLOADS[args.workload]
    if args.n:
        n = args.n
    q = p
    locs, X, y, L = make_inputs(n, p, seed_off)
    thetas = theta_stencil(n, p, q, L, y)
    K, Wm = args.steps, args.warmup
    obj = distributed_objective(y, X, locs, cov_name=cov, profileLik=True, device=local, nb_outer=args.nb)
    for s in range(Wm):
        obj.n2LL(thetas[s % len(thetas)])
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    dist.barrier(); torch.cuda.synchronize()
    l0 = obj.launch_count
    t0 = time.perf_counter()
    phase = {"build_ms": 0.0, "chol_ms": 0.0, "finalize_ms": 0.0}
    vals = []
    for s in range(K):
        vals.append(obj.n2LL(thetas[(Wm + s) % len(thetas)]))
        tm = obj.timings()
        for k in phase:
            phase[k] += tm[k]
    dist.barrier(); torch.cuda.synchronize()
    elapsed = time.perf_counter() - t0
    launches = obj.launch_count - l0
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([elapsed, float(launches)], dtype=torch.float64, device="cuda")
        tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        elapsed, launches = float(tmax[0]), int(t[1].item())
    if rank == 0:
        pr, pc = process_grid(world)
        if os.environ.get("VC_GRID"):
            pr, pc = (int(v) for v in os.environ["VC_GRID"].split("x"))
        chol_s = phase["chol_ms"] / K * 1e-3
        line = {"metric": "-2logL evals/sec (fp64), one evaluation sharded 2D block-cyclic", "value": K / elapsed,
                "unit": UNIT, "n_gpus": world, "steps": K, "warmup": Wm, "ms_per_step": elapsed / K * 1e3,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": desc, "n": n, "p": p, "q": q, "cov": cov, "grid": "%dx%d" % (pr, pc), "nb_outer": args.nb or "auto",
                           "parallelism": "2D block-cyclic Sigma_y, NCCL panel broadcasts, lookahead",
                           "l2": "inputs larger than L2"},
                "clocks": clocks, "gpu_launches": launches,
                "e2e": {"value": K / elapsed, "unit": UNIT, "h2d_bytes_per_step": 8 * (2 * q + 1), "d2h_bytes_per_step": 8 * 132,
                        "note": "inputs are O(n) and resident; theta in, scalar out through the public API"},
                "roofline": {"bound": "tensor", "kernel": "k_syrk_ws across %d GPUs" % world,
                             "achieved": n ** 3 / 3.0 / chol_s / 1e12, "peak": 37.0 * world, "unit": "TFLOP/s",
                             "frac": n ** 3 / 3.0 / chol_s / 1e12 / (37.0 * world), "traffic": None,
                             "peak_source": "N x 37.0 TFLOP/s fp64 DMMA probe (vc_probe_dmma_peak)"},
                "phases_ms": {k: v / K for k, v in phase.items()}, "n2ll_sample": vals[:2]}
        emit(line)
    obj.close()
    dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--n", "--problem-n", dest="n", type=int, default=0, help="override n (debug; use --problem-n under torchrun)")
    ap.add_argument("--cpu-sample-n", type=int, default=6000)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--nb", type=int, default=0, help="outer Cholesky block (0 = auto)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "c4":
        run_dist(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
