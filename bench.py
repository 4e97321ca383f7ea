#!/usr/bin/env python
"""bench.py -- headline benchmark of the -2 logL objective (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c3]

One "step" = one evaluation of -2 logL (fused covariance build -> bordered DMMA Cholesky -> fused
GLS / quadratic form / log-det) on synthetic sample_SVCdata-style inputs.  Default workload is the
headline configuration C3 of BASELINE.json: n = 50 000, p = q = 3, exponential covariance, fp64.

N > 1 (launched by torch.distributed.run, one rank per GPU): the headline line is the objective sharded
across the finite-difference stencil optim evaluates -- every rank holds a replica of the O(n) inputs
and evaluates its own theta; no collective on the data path ("scaling": "weak").  After the timed
region the same ranks also run the within-evaluation sharded path (C4: n = 150 000, p = q = 5, 2D
block-cyclic + NCCL panel broadcasts) and report it in the `sharded` record, with its parity against
the single-GPU path at n = 50 000.  `--workload c4` runs only that path as the headline.

Outside the timed region, at N = 1: `parity` (C3 result vs an independent cuSOLVER factorisation of the
GPU-built Sigma_y at n = 50 000), `library_bar` (cublasDgemm / cusolverDn potrf on this GPU),
`cpu_baseline` (the reference's CPU op sequence on the host cores, bounded sample).

`--impl reference` times the reference's CPU op sequence (numpy/LAPACK restatement in oracle/; the
reference itself is R and cannot run here) on all host cores.  A step is one faithful evaluation at a
bounded sample size n_s; `ms_per_step` is that measured time.  `value` (evals/s at the workload's n)
is extrapolated with an exponent fitted to measured samples and flagged as such; the lean in-place
sequence is measured once at the full n when host RAM allows (`cpu_baseline.lean_measured`).
"""
from __future__ import annotations

import os
import sys


def _host_threads() -> int:
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


# The CPU arm must use every host core: torch.distributed.run exports OMP_NUM_THREADS=1 to its workers, which
# silently made the N > 1 reference arm single-threaded.  Decide the BLAS thread count BEFORE numpy is imported.
_WORLD = int(os.environ.get("WORLD_SIZE", "1"))
if "reference" in sys.argv or _WORLD == 1:
    _t = str(int(os.environ.get("VC_CPU_THREADS", "0")) or _host_threads())
    for _k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[_k] = _t

import argparse
import json
import subprocess
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


def log(*a):
    print("[bench]", *a, file=sys.stderr, flush=True)


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
# DRAM bytes per evaluation from ncu (dram__bytes_read.sum + dram__bytes_write.sum), profiles/r02_launches_traffic_n50k.md
NCU_TRAFFIC_BYTES = {"c3": 6.157e11}           # potrf + trsm + syrk + strip kernels of one C3 evaluation
NCU_BUILD_TRAFFIC_BYTES = {"c3": 1.002e10}     # k_build_fast + k_border (algorithmic: 1.0e10 written)
# fp64-pipe instructions per element of k_build_fast<exp, d=2, q=3>, counted in its SASS (profiles/r02_build_alu_floor.md)
BUILD_FP64_INSTR_PER_ELEMENT = {"c3": 87.5}
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


def config_dict(workload, n, world):
    """The `config` object: identical for the GPU arm and the reference arm of the same run."""
    n0, p, cov, taper, _, desc = WORKLOADS[workload]
    n_pad = (n + 127) // 128 * 128
    return {"workload": desc, "n": n, "n_padded": n_pad, "p": p, "q": p, "cov": cov, "tapering": taper,
            "mu_mode": "profile/GLS", "theta": "optim central-difference stencil around default init, cycled",
            "parallelism": "theta-parallel replicas, no collective" if world > 1 else "single GPU",
            "l2": "inputs larger than L2: every evaluation rewrites and refactors the %.1f GB matrix" % (
                8.0 * n_pad * (n_pad + 128) / 1e9)}


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
# CPU arm: the reference's op sequence (oracle port) on the host cores
# ------------------------------------------------------------------------------------------------
def blas_threads():
    try:
        from threadpoolctl import threadpool_info
        return int(max([i.get("num_threads", 1) for i in threadpool_info()] + [1]))
    except Exception:
        return int(os.environ.get("OMP_NUM_THREADS", "1"))


def cpu_eval_timed(n_s, p, cov, seed_off, faithful=True):
    """One evaluation of the reference's dense n2LL op sequence at n_s on the host BLAS threads.
    Returns (t_build, t_cubic, t_dist, value): Sigma_y build (O(n^2)); chol + solves + determinant (O(n^3));
    the distance matrix (once per fit in the reference: not charged to an evaluation)."""
    from oracle import svc_oracle as o
    locs, X, y, L = make_inputs(n_s, p, seed_off)
    th = theta_stencil(n_s, p, p, L, y)[0]
    t0 = time.perf_counter()
    D = o.own_dist(locs)
    t_dist = time.perf_counter() - t0
    t0 = time.perf_counter()
    S = o.sigma_y(th, D, X, cov)
    t_build = time.perf_counter() - t0
    del S
    t0 = time.perf_counter()
    v = o.n2ll(th, D, X, X, y, cov, profile=True, faithful=faithful)
    t_all = time.perf_counter() - t0
    return t_build, max(t_all - t_build, 1e-9), t_dist, v


def fit_power(ns, ts):
    """least-squares fit of t = c n^e in log space"""
    ln, lt = np.log(np.asarray(ns, float)), np.log(np.asarray(ts, float))
    e, lc = np.polyfit(ln, lt, 1)
    return float(np.exp(lc)), float(e)


def extrapolate(samples, n):
    """samples: list of (n_s, t_build, t_cubic).  Build scales as n^2 (measured at the largest sample);
    the cubic part as c n^e with e the log-log slope between the two largest samples, clamped to [2, 3]
    (e = 3 when only one size was measured)."""
    samples = sorted(samples)
    n_l, tb_l, tc_l = samples[-1]
    if len({s[0] for s in samples}) >= 2:
        # local slope between the two largest samples (small sizes are dominated by thread start-up and the
        # single-threaded parts, which flattens a fit over all of them)
        c, e = fit_power([s[0] for s in samples[-2:]], [s[2] for s in samples[-2:]])
        e_used = min(max(e, 2.0), 3.0)      # asymptotically n^3; never extrapolate with a steeper or absurd slope
    else:
        e, e_used = 3.0, 3.0
    t_cubic = tc_l * (n / n_l) ** e_used
    t_build = tb_l * (n / n_l) ** 2
    return t_build + t_cubic, {"fitted_exponent": e, "exponent_used": e_used,
                               "samples": [{"n": s[0], "build_s": s[1], "cubic_s": s[2]} for s in samples]}


def full_size_cpu(n, p, cov, seed_off, threads, with_dgesv):
    """Measured at the FULL n on the host cores (needs 8 n^2 bytes of RAM, twice that with_dgesv):
    the lean in-place sequence, and -- with_dgesv -- one of the reference's three dgesv calls on t(R)."""
    from oracle.lean_inplace import n2ll_lean_inplace
    try:
        import psutil
        avail = psutil.virtual_memory().available
    except Exception:
        avail = 0
    need = 8.0 * n * n * 1.15 + 4e9
    if avail < need:
        return {"skipped": "host RAM available %.0f GB < %.0f GB needed" % (avail / 1e9, need / 1e9)}
    with_dgesv = with_dgesv and avail >= 2 * need
    locs, X, y, L = make_inputs(n, p, seed_off)
    th = theta_stencil(n, p, p, L, y)[0]
    t0 = time.perf_counter()
    r = n2ll_lean_inplace(th, locs, X, X, y, cov, threads=threads, time_dgesv=with_dgesv)
    wall = r["t_build"] + r["t_chol"] + r["t_solve"]
    out = {"n": n, "seconds": wall, "evals_per_s": 1.0 / wall, "build_s": r["t_build"], "dpotrf_s": r["t_chol"],
           "solve_s": r["t_solve"], "dpotrf_gflops": n ** 3 / 3.0 / r["t_chol"] / 1e9, "n2ll": r["n2ll"],
           "logdet": r["logdet"], "mu": [float(v) for v in r["mu"]], "theta": [float(v) for v in th],
           "sequence": "blocked in-place build (lower) -> dpotrf -> dtrtrs -> sum log diag; numpy/OpenBLAS, %d threads" % threads}
    if "t_dgesv" in r:
        out["dgesv_s"] = r["t_dgesv"]
        out["dgesv_gflops"] = 2.0 * n ** 3 / 3.0 / r["t_dgesv"] / 1e9
        out["dgesv_check"] = r["dgesv_check"]
    out["wall_s"] = time.perf_counter() - t0
    return out


def cpu_baseline_gpu_arm(n, p, cov, seed_off, threads):
    """cpu_baseline of the GPU arm (~10-40 s of CPU work): the lean in-place sequence -- the cheapest correct CPU
    evaluation, NOT what R executes -- measured at the full n when host RAM allows; else the faithful sequence
    at a bounded size, scaled with n^3."""
    full = full_size_cpu(n, p, cov, seed_off, threads, with_dgesv=False)
    if "skipped" not in full:
        return {"value": full["evals_per_s"], "unit": UNIT, "cores": threads, "kind": "port", "extrapolated": False,
                "sample": ("ONE full evaluation at n=%d, p=q=%d, %s with the LEAN in-place CPU sequence (blocked build of the "
                           "lower triangle, dpotrf, dtrtrs, sum log diag; oracle/lean_inplace.py): %.1f s = build %.1f + dpotrf "
                           "%.1f (%.0f GFLOP/s) + solves %.1f, %d BLAS threads.  The op sequence R actually executes (3 dgesv on "
                           "the triangular factor, (q+2) n x n temporaries) costs ~9x the flops: `bench.py --impl reference`"
                           % (n, p, cov, full["seconds"], full["build_s"], full["dpotrf_s"], full["dpotrf_gflops"],
                              full["solve_s"], threads)),
                "sample_seconds": full["seconds"], "n2ll": full["n2ll"], "theta": full["theta"], "detail": full}
    b, c, _, _ = cpu_eval_timed(8000, p, cov, seed_off, faithful=True)
    t_full, fit = extrapolate([(8000, b, c)], n)
    return {"value": 1.0 / t_full, "unit": UNIT, "cores": threads, "kind": "port", "extrapolated": True,
            "sample": "faithful R op sequence at n_s=8000 scaled with n^2 (build) and n^3 (cubic part); " + full["skipped"],
            "sample_seconds": b + c, "fit": fit}


def run_reference(args):
    n, p, cov, taper, seed_off, desc = WORKLOADS[args.workload]
    if args.n:
        n = args.n
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = blas_threads()
    n_s = args.cpu_sample_n
    t_start = time.perf_counter()
    for _ in range(max(0, args.warmup)):
        cpu_eval_timed(min(n_s, 2000), p, cov, seed_off)
    K = max(1, args.steps)
    tb, tc, step_s = [], [], []
    for _ in range(K):
        t0 = time.perf_counter()
        b, c, t_dist, _ = cpu_eval_timed(n_s, p, cov, seed_off, faithful=True)
        step_s.append(time.perf_counter() - t0 - t_dist)   # the distance matrix is per fit, not per evaluation
        tb.append(b); tc.append(c)
    samples = [(n_s, float(np.median(tb)), float(np.median(tc)))]
    for n_f in args.cpu_fit_n:
        if n_f and n_f != n_s:
            b, c, _, _ = cpu_eval_timed(n_f, p, cov, seed_off, faithful=True)
            samples.append((n_f, b, c))
            log("reference: faithful sample n=%d build %.2fs cubic %.2fs" % (n_f, b, c))
    t_extrap, fit = extrapolate(samples, n)
    n_l, tb_l, tc_l = max(samples)
    build_full = tb_l * (n / n_l) ** 2            # the O(n^2) numpy passes of the faithful Sigma_y loop
    full = None
    if not args.no_lean_full:
        full = run_child("cpu_full", args)          # 40 GB, 16 BLAS threads: isolated from this process
        if "failed" in full:
            full = {"skipped": full["failed"]}
        log("reference: measured at full n:", json.dumps(full)[:400])
    if full and "dgesv_s" in full:
        # cubic part of the faithful sequence MEASURED at the full n: dpotrf + 3 x dgesv(t(R), .) (one of the three
        # identical factorisations timed); only the O(n^2) build is scaled from the largest faithful sample
        t_full = build_full + full["dpotrf_s"] + 3.0 * full["dgesv_s"]
        how = ("cubic part MEASURED at n=%d: dpotrf %.1f s + 3 x dgesv on t(R) %.1f s (one of the three identical LU "
               "factorisations timed, %.0f GFLOP/s); Sigma_y build %.1f s scaled x (n/n_l)^2 from the faithful sample at "
               "n_l=%d" % (n, full["dpotrf_s"], full["dgesv_s"], full["dgesv_gflops"], build_full, n_l))
        extrapolated = "build part only (O(n^2)); the O(n^3) part is measured at the full n"
    else:
        t_full = t_extrap
        how = ("extrapolated from the largest faithful sample n_l=%d: build x (n/n_l)^2, cubic part x (n/n_l)^%.2f "
               "(log-log slope of the two largest samples %.3f, clamped to [2, 3])" % (n_l, fit["exponent_used"],
                                                                                      fit["fitted_exponent"]))
        extrapolated = True
    arm = {"value": 1.0 / t_full, "unit": UNIT, "cores": threads, "kind": "port", "extrapolated": extrapolated,
           "sample": ("the op sequence R executes (Sigma_y loop, dpotrf, 3 x dgesv on the triangular factor, log-det; "
                      "R-pkg/R/objective_functions.R:4-62, R-pkg/R/utils.R:93-101,171-194), %d BLAS threads.  Each timed step = "
                      "one such evaluation at the bounded size n_s=%d (median build %.3f s + cubic part %.3f s).  value: %s.  "
                      "The whole faithful sequence cannot run at n=%d ((q+5) 8 n^2 B = %.0f GB of temporaries)" % (
                          threads, n_s, samples[0][1], samples[0][2], how, n, (p + 5) * 8.0 * n * n / 1e9)),
           "sample_seconds": float(np.mean(step_s)), "seconds_per_eval_at_n": t_full, "fit": fit,
           "lean_measured": full}
    wall = time.perf_counter() - t_start
    line = {"impl": "reference", "metric": METRIC, "value": arm["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": float(np.mean(step_s)) * 1e3,
            "ms_per_step_note": "measured: one faithful evaluation at the bounded sample size n_s=%d; `value` is for n=%d "
                                "as described in cpu_baseline.sample" % (n_s, n),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": config_dict(args.workload, n, args.gpus),
            "cpu_baseline": arm,
            "e2e": {"value": arm["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "wall_s": wall}
    emit(line)


# ------------------------------------------------------------------------------------------------
# Records outside the timed region run in CHILD processes: they use other native stacks (cuSOLVER / cuBLAS through
# torch, OpenBLAS on 16 threads) on 20-60 GB, and a native fault in one of them must not take the headline line
# with it (one driver-like run of round 2 died with SIGSEGV after the timed legs; it did not reproduce).
# ------------------------------------------------------------------------------------------------
def run_child(kind, args, device=0, timeout=1500):
    cmd = [sys.executable, os.path.abspath(__file__), "--child", kind, "--workload", args.workload, "--device", str(device)]
    if args.n:
        cmd += ["--problem-n", str(args.n)]
    env = dict(os.environ)
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "MASTER_ADDR", "MASTER_PORT"):
        env.pop(k, None)
    try:
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=timeout, env=env)
    except subprocess.TimeoutExpired:
        return {"failed": "child %s timed out after %d s" % (kind, timeout)}
    lines = [ln for ln in r.stdout.splitlines() if ln.strip().startswith("{")]
    if r.returncode != 0 or not lines:
        return {"failed": "child %s exited with code %d" % (kind, r.returncode), "stderr_tail": r.stderr[-400:]}
    try:
        return json.loads(lines[-1])
    except Exception as e:  # noqa: BLE001
        return {"failed": "child %s printed no JSON: %r" % (kind, e)}


def child_main(args):
    """`bench.py --child KIND`: one record, printed as one JSON object on stdout."""
    n, p, cov, taper, seed_off, desc = WORKLOADS[args.workload]
    if args.n:
        n = args.n
    if args.child == "cpu_baseline":            # numpy / OpenBLAS only: torch is never imported here
        emit(cpu_baseline_gpu_arm(n, p, cov, seed_off, blas_threads()))
        return
    if args.child == "cpu_full":                # reference arm: lean sequence + one dgesv at the full n
        emit(full_size_cpu(n, p, cov, seed_off, blas_threads(), with_dgesv=True))
        return
    import torch
    from varycoef_b200 import SVCObjective
    torch.cuda.set_device(args.device)
    out = {}
    locs, X, y, L = make_inputs(n, p, seed_off)
    theta = theta_stencil(n, p, p, L, y)[0]
    if args.child in ("parity", "parity_library"):
        try:
            with SVCObjective(y, X, locs, cov_name=cov, tapering=taper, profileLik=True, device=args.device) as obj:
                out["parity"] = parity_record(obj, theta, X, y, args.device)
        except Exception as e:  # noqa: BLE001
            out["parity"] = {"failed": repr(e)}
        torch.cuda.empty_cache()
    if args.child in ("library_bar", "parity_library"):
        try:
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            from third_path import library_bar
            out["library_bar"] = library_bar(n, args.device)
        except Exception as e:  # noqa: BLE001
            out["library_bar"] = {"failed": repr(e)}
    emit(out)


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def parity_record(obj, theta, X, y, local):
    """C3 result vs an independent cuSOLVER/cuBLAS evaluation on the GPU-built Sigma_y (tools/third_path.py)."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    from third_path import third_path_n2ll
    ours = obj.n2LL(theta, parts=True)
    ref = third_path_n2ll(obj, theta, X, y, local)
    rel = abs(ours["n2ll"] - ref["n2ll"]) / abs(ref["n2ll"])
    return {"against": "cuSOLVER potrf + cuBLAS trsm of the GPU-built Sigma_y (tools/third_path.py): independent "
                       "factorisation, solves, GLS and log-det; the build kernel is shared (element-wise oracle tests)",
            "n": obj.n, "theta": [float(v) for v in theta], "n2ll": ours["n2ll"], "n2ll_third_path": ref["n2ll"],
            "rel": rel, "logdet_rel": abs(ours["logdet"] - ref["logdet"]) / abs(ref["logdet"]),
            "quad_rel": abs(ours["quad"] - ref["quad"]) / abs(ref["quad"]),
            "mu_rel": float(np.max(np.abs(ours["mu"] - ref["mu"])) / np.max(np.abs(ref["mu"]))),
            "tolerance": 1e-10, "ok": bool(rel <= 1e-10),
            "third_path_potrf_ms": ref["potrf_ms"]}


def sharded_record(args, dist, world, rank, local, peak_tf, single_c3):
    """Within-evaluation sharded path on the ranks of this run: C4 (n = 150 000, p = q = 5) timed, and its
    parity against the single-GPU path at C3 (single_c3 = (theta, n2ll, logdet) computed before)."""
    import torch
    from varycoef_b200.dist import distributed_objective, process_grid
    pr, pc = process_grid(world)
    rec = {"grid": "%dx%d" % (pr, pc)}
    # parity first (small): the same N-GPU driver at C3 vs the single-GPU value
    n3, p3, cov3, _, so3, _ = WORKLOADS["c3"]
    locs, X, y, L = make_inputs(n3, p3, so3)
    theta3, v_single, ld_single = single_c3
    obj = distributed_objective(y, X, locs, cov_name=cov3, profileLik=True, device=local)
    r = obj.n2LL(theta3, parts=True)
    obj.close()
    rec["parity_n"] = n3
    rec["parity_n2ll_sharded"] = r["n2ll"]
    rec["parity_n2ll_single_gpu"] = v_single
    rec["parity_rel"] = abs(r["n2ll"] - v_single) / abs(v_single)
    rec["parity_logdet_rel"] = abs(r["logdet"] - ld_single) / abs(ld_single)
    del locs, X, y
    # C4 timed
    n, p, cov, _, so, desc = WORKLOADS["c4"]
    if args.sharded_n:
        n = args.sharded_n
    locs, X, y, L = make_inputs(n, p, so)
    thetas = theta_stencil(n, p, p, L, y)
    obj = distributed_objective(y, X, locs, cov_name=cov, profileLik=True, device=local, nb_outer=args.nb)
    Ks, Ws = args.sharded_steps, 1
    for s in range(Ws):
        obj.n2LL(thetas[s % len(thetas)])
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    dist.barrier(); torch.cuda.synchronize()
    l0 = obj.launch_count
    t0 = time.perf_counter()
    phase = {"build_ms": 0.0, "chol_ms": 0.0, "finalize_ms": 0.0}
    vals = []
    for s in range(Ks):
        vals.append(obj.n2LL(thetas[(Ws + s) % len(thetas)]))
        tm = obj.timings()
        for k in phase:
            phase[k] += tm[k]
    dist.barrier(); torch.cuda.synchronize()
    elapsed = time.perf_counter() - t0
    launches = obj.launch_count - l0
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([elapsed, float(launches)], dtype=torch.float64, device="cuda")
    tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    elapsed, launches = float(tmax[0]), int(t[1].item())
    rx, bound = obj.dist_traffic()
    obj.close()
    chol_s = phase["chol_ms"] / Ks * 1e-3
    rec.update({"nvlink_rx_bytes_per_rank": rx, "nvlink_rx_gbs_per_rank": rx / chol_s / 1e9,
                "nvlink_note": "panel + L_DD broadcasts on the world communicator: every rank receives the whole panel "
                               "(~4 n^2 B); a row/column-communicator scheme would need nvlink_bytes_per_rank_bound"})
    rec.update({"workload": desc, "n": n, "p": p, "q": p, "cov": cov, "steps": Ks, "warmup": Ws,
                "evals_per_s": Ks / elapsed, "ms_per_eval": elapsed / Ks * 1e3, "scaling": "strong",
                "tflops": n ** 3 / 3.0 / chol_s / 1e12, "peak_tflops": peak_tf * world,
                "frac": n ** 3 / 3.0 / chol_s / 1e12 / (peak_tf * world) if peak_tf > 0 else None,
                "peak_source": "N x fp64 DMMA probe of rank 0 (vc_probe_dmma_peak)",
                "phases_ms": {k: v / Ks for k, v in phase.items()}, "gpu_launches": launches, "clocks": clocks,
                "n2ll_sample": vals[:2],
                "nvlink_bytes_per_rank_bound": bound,
                "parallelism": "2D block-cyclic Sigma_y, NCCL panel broadcasts, lookahead"})
    return rec


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

    # ---- records outside the timed region ----
    extra = {}
    at_size = (n == WORKLOADS[args.workload][0])
    single_c3 = None
    if world > 1 and not args.no_sharded and args.workload == "c3" and at_size:
        r0 = obj.n2LL(thetas[0], parts=True)
        single_c3 = (thetas[0], r0["n2ll"], r0["logdet"])
    n_pad = obj.padded_n
    obj.close()
    del obj
    torch.cuda.empty_cache()
    if world == 1 and not (args.no_parity and args.no_library_bar):
        kind = "parity_library" if not (args.no_parity or args.no_library_bar) else ("parity" if args.no_library_bar else "library_bar")
        rec = run_child(kind, args, device=local)
        if "failed" in rec:
            extra["parity" if kind != "library_bar" else "library_bar"] = rec
        else:
            extra.update(rec)
    if single_c3 is not None:
        try:
            extra["sharded"] = sharded_record(args, dist, world, rank, local, peak_tf.value, single_c3)
        except Exception as e:  # noqa: BLE001
            extra["sharded"] = {"failed": repr(e)}
            log("sharded record failed on rank %d: %r" % (rank, e))
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

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
            "traffic": NCU_TRAFFIC_BYTES.get(args.workload) if at_size else None,
            "traffic_note": "DRAM read + write bytes of all Cholesky kernels of ONE evaluation (per evaluation, like "
                            "`achieved`), ncu dram__bytes_{read,write}.sum, profiles/r02_launches_traffic_n50k.md: 616 GB in "
                            "1.21 s = 8 % of DRAM peak -- the factorisation is DMMA-bound (pipe 95.7 % active in "
                            "k_syrk_ws, profiles/r01_syrk_ws_n20k.md); the C trapezoid alone accounts for ~400 GB",
            "peak_source": "fp64 DMMA m8n8k4 register-resident probe measured live on this GPU (vc_probe_dmma_peak); "
                           "MEASURED_PEAKS.json has no fp64 entry; see library_bar.dgemm_tflops for cuBLAS on the same GPU",
            "algorithmic": "n^3/3 flops per evaluation / CUDA-event time of the whole factorisation (panels included)"}
    roof_build = {"bound": "hbm", "kernel": "k_build_fast (fused distance + covariance)", "achieved": build_bytes / build_s / 1e9,
                  "peak": hbm_peak, "unit": "GB/s", "frac": build_bytes / build_s / 1e9 / hbm_peak,
                  "traffic": NCU_BUILD_TRAFFIC_BYTES.get(args.workload) if at_size else None, "peak_source": hbm_src,
                  "algorithmic": "4 n (n+1) bytes written (lower triangle) / CUDA-event time of build + border"}
    ipe = BUILD_FP64_INSTR_PER_ELEMENT.get(args.workload) if at_size else None
    if ipe and clocks and clocks.get("sm_mhz"):
        # the kernel is bound by the fp64 pipe, not by HBM: floor = elements x instructions / (SMs x 64 lanes x clock)
        floor_s = (n * (n + 1) / 2.0) * ipe / (148 * 64 * clocks["sm_mhz"] * 1e6)
        roof_build["alu_floor"] = {"fp64_instr_per_element": ipe, "floor_ms": floor_s * 1e3, "frac_of_floor": floor_s / build_s,
                                   "source": "SASS count of the inner loop of k_build_fast<exp,2,3> (profiles/r02_build_alu_floor.md): "
                                             "DFMA+DMUL+DADD+DSETP per element at 64 fp64 lanes/clk/SM x 148 SMs"}
    h2d = 8 * n * (2 + p + q + 1) + 8 * (2 * q + 1)
    d2h = 8 * 132
    line = {
        "metric": METRIC, "value": K * world / elapsed, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": Wm,
        "ms_per_step": elapsed / K * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": config_dict(args.workload, n, world),
        "clocks": clocks,
        "e2e": {"value": K * world / elapsed_e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": launches_total,
        "roofline": roof, "roofline_build": roof_build,
        "phases_ms": {k: v / K for k, v in phase.items()},
        # device time of one step from CUDA events recorded on the library's own stream around build / Cholesky /
        # finalize (rank 0; `ms_per_step` is the wall clock of the synchronised region, max over ranks)
        "ms_per_step_device": sum(phase.values()) / K,
        "n2ll_sample": vals[:2],
    }
    line.update(extra)
    if world == 1 and not args.no_cpu:
        line["cpu_baseline"] = run_child("cpu_baseline", args)
        if "failed" in line["cpu_baseline"]:
            line["cpu_baseline"].update({"value": None, "unit": UNIT, "cores": _host_threads(), "kind": "port", "sample": "child process failed"})
    emit(line)
    if dist is not None:
        dist.destroy_process_group()


def run_dist(args):
    """--workload c4: ONE evaluation sharded 2D block-cyclic over all ranks (n = 150 000 does not fit
    one GPU in full storage).  value = whole-job evaluations per second; strong scaling in N."""
    import torch
    import torch.distributed as dist
    import ctypes as C
    from varycoef_b200 import _lib
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
    peak_tf = C.c_double(0.0)
    _lib.lib().vc_probe_dmma_peak(local, 20000, C.byref(peak_tf))
    if world == 1 and not args.force_dist:
        # lower block-column storage: n = 150 000 (91 GB) fits ONE B200 -- the plain single-GPU path is the N = 1
        # point of the strong-scaling curve and the bit-exact parity reference of the sharded runs
        from varycoef_b200 import SVCObjective
        obj = SVCObjective(y, X, locs, cov_name=cov, profileLik=True, device=local, nb_outer=args.nb)
    else:
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
        peak = peak_tf.value * world
        line = {"metric": "-2logL evals/sec (fp64), one evaluation sharded 2D block-cyclic", "value": K / elapsed,
                "unit": UNIT, "n_gpus": world, "steps": K, "warmup": Wm, "ms_per_step": elapsed / K * 1e3,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": desc, "n": n, "p": p, "q": q, "cov": cov, "grid": "%dx%d" % (pr, pc), "nb_outer": args.nb or "auto",
                           "parallelism": ("2D block-cyclic Sigma_y, NCCL panel broadcasts, lookahead" if world > 1 or args.force_dist
                                           else "single GPU (plain path, lower block-column storage)"),
                           "l2": "inputs larger than L2"},
                "clocks": clocks, "gpu_launches": launches,
                "e2e": {"value": K / elapsed, "unit": UNIT, "h2d_bytes_per_step": 8 * (2 * q + 1), "d2h_bytes_per_step": 8 * 132,
                        "note": "inputs are O(n) and resident; theta in, scalar out through the public API"},
                "roofline": {"bound": "tensor", "kernel": "k_syrk_ws across %d GPUs" % world,
                             "achieved": n ** 3 / 3.0 / chol_s / 1e12, "peak": peak, "unit": "TFLOP/s",
                             "frac": n ** 3 / 3.0 / chol_s / 1e12 / peak if peak > 0 else None, "traffic": None,
                             "peak_source": "N x fp64 DMMA probe of rank 0 (vc_probe_dmma_peak)"},
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
    ap.add_argument("--cpu-sample-n", type=int, default=4000, help="reference arm: size of one step's faithful evaluation")
    ap.add_argument("--cpu-fit-n", type=lambda s: [int(v) for v in s.split(",") if v], default=[12000],
                    help="reference arm: extra faithful sample sizes (once each) for the fitted exponent")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-lean-full", action="store_true", help="reference arm: skip the lean in-place run at full n")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-library-bar", action="store_true")
    ap.add_argument("--no-sharded", action="store_true", help="N > 1: skip the sharded (C4) record")
    ap.add_argument("--sharded-n", type=int, default=0, help="override n of the sharded record (debug)")
    ap.add_argument("--sharded-steps", type=int, default=2)
    ap.add_argument("--nb", type=int, default=0, help="outer Cholesky block (0 = auto)")
    ap.add_argument("--child", default="", help=argparse.SUPPRESS)     # internal: one record in a child process
    ap.add_argument("--device", type=int, default=0, help=argparse.SUPPRESS)
    ap.add_argument("--force-dist", action="store_true", help="--workload c4 at N = 1: use the block-cyclic driver on a 1x1 grid")
    args = ap.parse_args()
    if args.child:
        child_main(args)
    elif args.impl == "reference":
        run_reference(args)
    elif args.workload == "c4":
        run_dist(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
