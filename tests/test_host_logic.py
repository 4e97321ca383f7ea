"""CPU-side checks: the C-ABI library loads and exports every declared symbol, the host mirror
of the reference interface behaves like the reference, and the product never touches oracle/."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "varycoef_b200.h")).read()
    return sorted(set(re.findall(r"VC_API\s+[\w\s\*]+?\b(vc_\w+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from varycoef_b200 import _lib
    L = _lib.lib()
    syms = _header_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(L, s), "symbol %s declared in include/varycoef_b200.h is not exported" % s
        assert s in _lib.SIGNATURES, "no ctypes signature for %s" % s
    assert sorted(_lib.SIGNATURES) == syms
    assert L.vc_abi_version() == 1


def test_config_default_and_struct_layout():
    from varycoef_b200 import _lib
    cfg = _lib.VcConfig()
    assert _lib.lib().vc_config_default(C.byref(cfg)) == 0
    assert (cfg.cov_id, cfg.dist_method, cfg.taper_id, cfg.nb_outer, cfg.flags) == (0, 0, 0, 0, 0)
    assert cfg.minkowski_p == 2.0
    assert C.sizeof(_lib.VcConfig) == 40 and C.sizeof(_lib.VcTiming) == 16


def test_block_cyclic_maps():
    L = __import__("varycoef_b200._lib", fromlist=["lib"]).lib()
    pr, pc = 2, 4
    seen = {}
    for bi in range(7):
        for bj in range(9):
            r = L.vc_bc_owner(bi, bj, pr, pc)
            assert r == (bi % pr) * pc + (bj % pc)
            seen.setdefault(r, 0)
            seen[r] += 1
    assert set(seen) == set(range(8))
    assert [L.vc_bc_local(b, 4) for b in range(9)] == [0, 0, 0, 0, 1, 1, 1, 1, 2]
    for nblocks in (1, 5, 8, 13):
        assert sum(L.vc_bc_count(nblocks, c, 4) for c in range(4)) == nblocks
    assert L.vc_bc_owner(-1, 0, 2, 2) == -1


@pytest.mark.skipif(__import__("torch").cuda.is_available(), reason="checks the no-GPU failure mode")
def test_product_fails_loudly_without_gpu():
    from varycoef_b200 import SVCObjective, _lib
    with pytest.raises(_lib.VcError):
        SVCObjective(np.zeros(10), np.ones((10, 2)), np.arange(10.0))


def test_invalid_arguments_rejected_before_touching_the_gpu():
    from varycoef_b200 import SVCObjective
    with pytest.raises(ValueError):
        SVCObjective(np.zeros(10), np.ones((10, 2)), np.arange(10.0), cov_name="gauss")
    with pytest.raises(ValueError):
        SVCObjective(np.zeros(10), np.ones((9, 2)), np.arange(10.0))
    with pytest.raises(ValueError):   # get_taper returns NULL for sph (R-pkg/R/utils.R:545-552)
        SVCObjective(np.zeros(10), np.ones((10, 2)), np.arange(10.0), cov_name="sph", tapering=1.0)
    with pytest.raises(ValueError):   # R-pkg/R/utils.R:211 + :3 : mat32 + taper needs q = 1
        SVCObjective(np.zeros(10), np.ones((10, 2)), np.arange(10.0), cov_name="mat32", tapering=1.0)
    with pytest.raises(ValueError):
        SVCObjective(np.zeros(10), np.ones((10, 2)), np.arange(10.0), dist={"method": "canberra"})


def test_no_oracle_or_cpu_fallback_in_product():
    pkg = os.path.join(ROOT, "varycoef_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "svc_oracle" not in src, f
    out = subprocess.run([sys.executable, "-c",
                          "import sys, varycoef_b200, varycoef_b200.svc_mle, varycoef_b200.optim;"
                          "print(any(m == 'oracle' or m.startswith('oracle.') for m in sys.modules))"],
                         cwd=ROOT, capture_output=True, text=True)
    assert out.stdout.strip() == "False", out.stderr


def test_pc_penalty_matches_oracle():
    from oracle import svc_oracle as o
    from varycoef_b200 import pc_penalty
    x = np.array([0.7, 1.2, 2.0, 0.4, 0.3])
    pcp = (0.5, 0.05, 2.0, 0.1)
    assert pc_penalty(x, 2, None) == 0.0
    assert pc_penalty(x, 2, pcp) == pytest.approx(o.pc_penalty(x, 2, pcp), rel=1e-15)
    lam_r = -np.log(0.05) * 2 * 0.5
    lam_s = -np.log(0.1) / 2.0
    ref = sum(4 * np.log(r) + lam_r / r + 2 * lam_s * np.sqrt(s) for r, s in ((0.7, 1.2), (2.0, 0.4)))
    assert pc_penalty(x, 2, pcp) == pytest.approx(ref, rel=1e-14)


def test_check_cov_lower_and_init_bounds():
    from oracle import svc_oracle as o
    from varycoef_b200 import check_cov_lower
    from varycoef_b200.svc_mle import SVC_mle_control, init_bounds_optim
    assert check_cov_lower([0.1, 0, 0.2, 1, 0.2], 2)
    assert not check_cov_lower([0, 0, 0.2, 1, 0.2], 2)
    assert not check_cov_lower([0.1, 0, 0.2, 1, 0], 2)
    assert not check_cov_lower([0.1, 0, 0.2, -1, 0], 2)
    for prof in (False, True):
        ctl = SVC_mle_control(profileLik=prof)
        liu = init_bounds_optim(ctl, 3, 2, range(5 if prof else 8), 2.5, 4.0, [1.0, 2.0, 3.0])
        ref = o.init_bounds_optim(3, 2, 2.5, 4.0, [1.0, 2.0, 3.0], prof)
        for k in ("lower", "init", "upper"):
            assert np.array_equal(liu[k], ref[k])
    ctl = SVC_mle_control(init=[1, 1, 1])
    with pytest.raises(ValueError, match="Initial vector has wrong length"):
        init_bounds_optim(ctl, 2, 2, range(7), 1.0, 1.0, [0, 0])
    ctl = SVC_mle_control(lower=[0.0, 0, 0.1, 0, 0.1, 0, 0], profileLik=False)
    with pytest.raises(ValueError, match="Lower boundary vector is not greater"):
        init_bounds_optim(ctl, 2, 2, range(7), 1.0, 1.0, [0, 0])
    with pytest.raises(ValueError):
        SVC_mle_control(cov_name="gauss")
    with pytest.raises(ValueError):
        SVC_mle_control(tapering=-1.0)
    assert SVC_mle_control()["profileLik"] is False and SVC_mle_control()["hessian"] is True


def test_median_dist_exact_small_and_histogram_paths():
    from oracle.svc_oracle import median_dist
    from oracle import svc_oracle as o
    rng = np.random.default_rng(3)
    for n, d in ((7, 1), (50, 2), (501, 2)):
        locs = rng.uniform(0, 5, (n, d))
        assert median_dist(locs) == pytest.approx(float(np.median(o.own_dist(locs))), rel=1e-15)
    locs = rng.uniform(0, 5, (3200, 2))      # > 4e6 pairs: histogram selection path
    assert median_dist(locs, block=1024) == pytest.approx(float(np.median(o.own_dist(locs))), rel=1e-15)
    locs = rng.uniform(0, 5, (60, 3))
    for m in ("maximum", "manhattan"):
        assert median_dist(locs, m) == pytest.approx(float(np.median(o.own_dist(locs, method=m))), rel=1e-15)


def test_optim_driver_semantics():
    from varycoef_b200.optim import optim_lbfgsb, optimhess, fd_points
    # central differences clipped at the bounds, divisor = span actually used (R's fmingr)
    pts, div = fd_points(np.array([1.0, 0.0005]), np.array([2.0, 1.0]), np.array([1e-3, 1e-3]),
                         np.array([0.0, 0.0]), np.array([1.0, 1.0]))
    assert np.allclose(pts[0], [2.0, 0.0005]) and np.allclose(pts[1], [1.998, 0.0005])
    assert np.allclose(pts[3], [2.0, 0.0]) and np.allclose(div, [1e-3, 1.5e-3])
    # quadratic: exact Hessian, batch and scalar paths agree
    A = np.array([[3.0, 0.5], [0.5, 2.0]])
    f = lambda x: float(x @ A @ x)  # noqa: E731
    fb = lambda xs: np.array([f(x) for x in xs])  # noqa: E731
    H = optimhess(np.array([0.3, -0.2]), fn=f, parscale=np.array([2.0, 0.5]))
    assert np.allclose(H, 2 * A, rtol=1e-6)
    r1 = optim_lbfgsb([1.0, 1.0], fn=f, hessian=True)
    r2 = optim_lbfgsb([1.0, 1.0], fn_batch=fb, hessian=True)
    assert np.allclose(r1["par"], 0, atol=1e-4) and np.allclose(r1["par"], r2["par"])
    assert r1["convergence"] == 0 and r1["objective_evals"] == r2["objective_evals"]
    assert r1["objective_evals"] == r1["counts"]["function"] + 4 * r1["counts"]["gradient"] + 16
    # bound-constrained
    r = optim_lbfgsb([2.0, 2.0], fn=lambda x: float((x[0] + 1) ** 2 + (x[1] - 3) ** 2), lower=[0, 0], upper=[5, 2.5])
    assert np.allclose(r["par"], [0.0, 2.5], atol=1e-6)
    with pytest.raises(FloatingPointError, match="finite values"):
        optim_lbfgsb([1.0], fn=lambda x: float("nan"))


def test_fast_exp_matches_libm():
    """vc_math.cuh's table-driven exp(-t) (the build kernel's hot op), compiled for the host."""
    import tempfile
    src = r'''
#include "vc_math.cuh"
#include <stdio.h>
#include <stdlib.h>
int main(){ double tab[64]; vc_fill_exp2_table(tab); double mx=0; srand(7);
 for(long i=0;i<4000000;i++){ double u=rand()/(double)RAND_MAX; int c=i%4;
  double t = c==0? u*1e-3 : c==1? u*5 : c==2? u*100 : u*707.9;
  double a=vc_exp_neg(t,tab), b=exp(-t); double r=fabs(a-b)/b; if(r>mx) mx=r; }
 printf("%.3e %.17g %.17g\n", mx, vc_exp_neg(0.0,tab), vc_exp_neg(800.0,tab)); return 0; }
'''
    with tempfile.TemporaryDirectory() as td:
        cpp = os.path.join(td, "t.cpp")
        open(cpp, "w").write(src)
        exe = os.path.join(td, "t")
        subprocess.check_call(["g++", "-O2", "-x", "c++", "-I", os.path.join(ROOT, "varycoef_b200", "csrc"), cpp, "-o", exe])
        mx, e0, e800 = subprocess.check_output([exe], text=True).split()
    assert float(mx) < 4.5e-16 and float(e0) == 1.0 and float(e800) == 0.0


def test_graft_entry_build_and_reference_arm_json():
    """build() compiles / loads the library without a GPU; `bench.py --impl reference` prints exactly one
    JSON line on stdout with the contract's keys."""
    import json
    sys.path.insert(0, ROOT)
    import __graft_entry__ as g
    g.build()
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "1", "--cpu-sample-n", "600", "--cpu-fit-n", "300,900", "--problem-n", "2500"],
                         cwd=ROOT, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-500:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "evals/s" and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["value"] > 0
    assert d["metric"].startswith("-2logL evals/sec")
    # ms_per_step is the measured time of one bounded step, not the extrapolated evaluation
    assert d["ms_per_step"] < 60e3 and d["cpu_baseline"]["extrapolated"]
    assert len(d["cpu_baseline"]["fit"]["samples"]) == 3
    # the lean in-place sequence was really run at the full (here: overridden) n and agrees with the oracle port
    lean = d["cpu_baseline"]["lean_measured"]
    assert lean["n"] == 2500 and lean["seconds"] > 0 and np.isfinite(lean["n2ll"])
    # one of the reference's three dgesv calls on t(R) was timed at the full n and reproduces the dtrtrs result
    assert lean["dgesv_s"] > 0 and lean["dgesv_check"] < 1e-10
    assert abs(1.0 / d["value"] - d["cpu_baseline"]["seconds_per_eval_at_n"]) < 1e-9 / d["value"]
    # same config keys as the GPU arm emits (bench.config_dict)
    for k in ("workload", "n", "n_padded", "p", "q", "cov", "tapering", "mu_mode", "theta", "parallelism", "l2"):
        assert k in d["config"], k


def test_r_shim_type_checks_against_the_header():
    """The .Call shim a varycoef maintainer would add must at least type-check against
    include/varycoef_b200.h; R itself is not in the image, r-shim/check/ holds a minimal stand-in of its C API."""
    out = subprocess.run(["gcc", "-std=c99", "-fsyntax-only", "-I" + os.path.join(ROOT, "r-shim", "check"),
                          "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "r-shim", "src", "varycoef_b200_r.c")],
                         capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    src = open(os.path.join(ROOT, "r-shim", "src", "varycoef_b200_r.c")).read()
    for fn in ("vc_create", "vc_create_multi", "vc_median_dist", "vc_n2ll", "vc_n2ll_batch", "vc_post", "vc_predict", "vc_get_whitened", "vc_gls_chol",
               "vc_destroy", "vc_last_error", "vc_last_info"):
        assert fn + "(" in src, fn


def test_plain_c_client_links_and_validates_arguments(tmp_path):
    """The drop-in boundary is a C ABI: a C99 program with no Python / torch in sight links against the library,
    gets VC_ERR_INVALID for bad arguments and -- in this GPU-less container -- a loud VC_ERR_CUDA from vc_create."""
    from varycoef_b200 import _build
    lib = _build.build()
    exe = str(tmp_path / "abi_client")
    out = subprocess.run(["gcc", "-std=c99", "-I" + os.path.join(ROOT, "include"),
                          os.path.join(ROOT, "tests", "c", "abi_client.c"), "-o", exe,
                          "-L" + os.path.dirname(lib), "-lvarycoef_b200", "-lm",
                          "-Wl,-rpath," + os.path.dirname(lib)], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    run = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert run.returncode == 0, (run.returncode, run.stdout, run.stderr)
    assert run.stdout.startswith("abi ")
    assert ("no gpu:" in run.stdout) or ("n2ll " in run.stdout)


def test_owner_major_panel_map_is_a_bijection_with_contiguous_owner_ranges():
    """Sharded path: the gathered panel buffer holds the row tiles >= oe in OWNER-major order (VcPanelMap,
    csrc/vc_common.cuh; filled by fill_panel_map in csrc/vc_dist.cu) so that every process row's part is ONE
    contiguous range = one NCCL broadcast.  vc_bc_panel_layout exposes the map the plan uses; the device function
    vc_panel_pos is additionally compiled for the host from the header.  For every step of a few grids: positions
    are a bijection onto 0 .. count-1, every owner's tiles are contiguous and in tile order, border tiles follow
    process row 0's main tiles, and header and library agree."""
    import ctypes as C
    import tempfile
    from varycoef_b200 import _lib
    L = _lib.lib()
    src = r"""
#include <stdio.h>
#include <string.h>
#include "vc_common.cuh"
int main(int argc, char** argv) {
  int nt, nb, pr, nbt, oe;
  while (scanf("%d %d %d %d %d", &nt, &nb, &pr, &nbt, &oe) == 5) {
    VcPanelMap m; memset(&m, 0, sizeof(m));
    m.pr = pr; m.nb = nb; m.nt = nt;
    const int nblocks = (nt + nb - 1) / nb, B0 = (oe + nb - 1) / nb;
    int pos = 0;
    for (int r = 0; r < pr; ++r) {                        // an independent restatement of the rule
      int bf = B0; while (bf % pr != r) ++bf;
      m.bfirst[r] = bf; m.chunk_off[r] = pos;
      int cnt = 0;
      for (int B = bf; B < nblocks; B += pr) cnt += (nt - B * nb < nb ? nt - B * nb : nb);
      if (r == 0) { m.main0 = cnt; cnt += nbt; }
      pos += cnt;
    }
    for (int I = oe; I < nt + nbt; ++I) printf("%d ", vc_panel_pos(m, I, oe));
    printf("\n");
  }
  return 0;
}
"""
    cases = []
    for nt, nb, pr, nbt in [(40, 4, 2, 1), (41, 4, 4, 1), (37, 8, 4, 2), (16, 2, 1, 1), (1172, 8, 4, 1)]:
        for oe in range(nb, nt, nb * (1 if nt < 100 else 37)):
            cases.append((nt, nb, pr, nbt, oe))
    with tempfile.TemporaryDirectory() as td:
        cpp = os.path.join(td, "t.cpp")
        open(cpp, "w").write(src)
        exe = os.path.join(td, "t")
        subprocess.check_call(["g++", "-O1", "-x", "c++", "-I", os.path.join(ROOT, "varycoef_b200", "csrc"),
                               "-I", "/usr/local/cuda/include", cpp, "-o", exe])
        out = subprocess.run([exe], input="".join("%d %d %d %d %d\n" % c for c in cases), capture_output=True, text=True,
                             check=True).stdout.strip().split("\n")
    assert len(out) == len(cases)
    for (nt, nb, pr, nbt, oe), line in zip(cases, out):
        pos = [int(v) for v in line.split()]
        tiles = list(range(oe, nt + nbt))
        off = (C.c_int32 * 8)()
        cnt = (C.c_int32 * 8)()
        lp = (C.c_int32 * len(tiles))()
        assert L.vc_bc_panel_layout(nt, nbt, nb, pr, oe, off, cnt, lp) == len(tiles)
        assert list(lp) == pos                                                        # library == header restatement
        assert sorted(pos) == list(range(len(tiles))), (nt, nb, pr, oe)
        owner = [((I // nb) % pr) if I < nt else 0 for I in tiles]
        for r in range(pr):
            qs = [q for q, o in zip(pos, owner) if o == r]
            assert len(qs) == cnt[r]
            if qs:
                assert qs == list(range(off[r], off[r] + cnt[r])), (nt, nb, pr, oe, r)  # contiguous, in tile order
    assert L.vc_bc_panel_layout(10, 1, 4, 9, 4, None, None, None) == -1


def test_batch_plan_shares_slots_between_repeated_thetas_and_puts_the_last_evaluation_in_slot_0():
    """vc_plan_batch (host part of vc_n2ll_batch): every distinct theta gets exactly one slot per call, evaluations
    that repeat it share that slot (one factorisation, several finalize jobs), groups hold at most S slots, the
    theta of the LAST valid evaluation sits in slot 0 of the LAST group (what vc_get_chol / vc_post refer to),
    result records are dense and in evaluation order within a group, invalid evaluations get nothing."""
    import ctypes as C
    from varycoef_b200 import _lib
    L = _lib.lib()
    rng = np.random.default_rng(0)
    for m, nd, S in [(1, 1, 1), (9, 4, 32), (21, 15, 4), (100, 41, 32), (50, 50, 7), (13, 1, 3)]:
        base = rng.uniform(0.5, 2.0, (nd, 7))
        pick = rng.integers(0, nd, m)
        pick[:nd] = np.arange(nd)[:m]                      # every distinct theta occurs (when m >= nd)
        thetas = np.ascontiguousarray(base[pick])
        valid = np.ones(m, dtype=np.int32)
        if m > 5:
            valid[rng.integers(0, m, 2)] = 0
        g, s, r = (np.empty(m, dtype=np.int32) for _ in range(3))
        ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int32))
        nu = L.vc_plan_batch(m, 7, _lib.ptr(thetas), ip(valid), S, ip(g), ip(s), ip(r))
        ok = valid == 1
        assert nu == len({tuple(t) for t in thetas[ok]})
        assert np.all(g[~ok] == -1) and np.all(s[~ok] == -1) and np.all(r[~ok] == -1)
        # same theta <-> same (group, slot)
        key = {}
        for e in np.nonzero(ok)[0]:
            key.setdefault(tuple(thetas[e]), set()).add((int(g[e]), int(s[e])))
        assert all(len(v) == 1 for v in key.values())
        assert len({next(iter(v)) for v in key.values()}) == nu           # distinct thetas never share a slot
        ngroups = int(g[ok].max()) + 1
        assert ngroups == -(-nu // S)
        for gi in range(ngroups):
            slots = sorted({int(x) for x in s[ok & (g == gi)]})
            assert slots == list(range(len(slots))) and len(slots) <= S   # dense slots 0 .. gsz-1
        last = int(np.nonzero(ok)[0][-1])
        assert g[last] == ngroups - 1 and s[last] == 0
        # records: dense, grouped by group, in evaluation order inside a group
        order = sorted(np.nonzero(ok)[0], key=lambda e: (g[e], e))
        assert [int(r[e]) for e in order] == list(range(len(order)))


def test_lower_block_column_storage_table():
    """Lower block-column storage: every block column stores the row tiles from its own
    diagonal block down, with its own leading dimension (an odd number of 128-double tiles).  The ranges must not
    overlap, must hold every lower tile, and one evaluation slot must be about half the full square:
    n = 50 000 -> 10.3 GB, n = 150 000 -> 91 GB (fits one B200)."""
    import ctypes as C
    from varycoef_b200 import _lib
    L = _lib.lib()
    for n, nb in [(1, 0), (127, 0), (1000, 0), (1000, 128), (5300, 384), (50000, 0), (150000, 0)]:
        cap = 4096
        off = (C.c_int64 * cap)()
        ld = (C.c_int32 * cap)()
        r0 = (C.c_int32 * cap)()
        tot = C.c_int64()
        nbc = L.vc_layout_describe(n, nb, C.byref(tot), off, ld, r0, cap)
        n_pad = -(-n // 128) * 128
        nt = n_pad // 128
        blk = (nb or (1024 if n_pad >= 40000 else (256 if n_pad <= 4096 else 512))) // 128
        assert nbc == -(-nt // blk)
        end_prev = 0
        for c in range(nbc):
            assert r0[c] == c * blk and off[c] == end_prev          # packed back to back, starts at its diagonal block
            rows = nt - r0[c]
            assert ld[c] % 128 == 0 and (ld[c] // 128) % 2 == 1 and ld[c] // 128 >= rows
            assert ld[c] // 128 <= rows + 1                          # at most one padding tile
            ncols = min(blk, nt - c * blk)
            end_prev = off[c] + ld[c] * ncols * 128
        assert tot.value == end_prev
        full = 8.0 * n_pad * (n_pad + 128)
        if n >= 50000:
            assert tot.value * 8 < 0.53 * full
    L.vc_layout_describe(50000, 0, C.byref(tot), None, None, None, 0)
    assert 10.0e9 < tot.value * 8 < 10.6e9
    L.vc_layout_describe(150000, 0, C.byref(tot), None, None, None, 0)
    assert 90e9 < tot.value * 8 < 92.5e9


def test_bench_child_records_fail_soft():
    """bench.py runs parity / library_bar / cpu_baseline in child processes so that a native fault there cannot take
    the headline JSON line with it.  In this GPU-less container the GPU child dies (no driver): the parent must get a
    {"failed": ...} record, not an exception; the CPU child must deliver a measured record."""
    code = r"""
import argparse, json, sys
sys.argv = ["bench.py"]
import bench
a = argparse.Namespace(workload="c3", n=1500)
gpu = bench.run_child("parity_library", a, device=0, timeout=300)
cpu = bench.run_child("cpu_baseline", a, timeout=300)
bench.emit({"gpu": gpu, "cpu": cpu})
"""
    out = subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-800:]
    import json
    d = json.loads([ln for ln in out.stdout.splitlines() if ln.startswith("{")][-1])
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if not has_gpu:
        assert "failed" in d["gpu"] and "exited with code" in d["gpu"]["failed"]
    assert d["cpu"]["kind"] == "port" and d["cpu"]["value"] > 0 and d["cpu"]["extrapolated"] is False
    assert d["cpu"]["detail"]["n"] == 1500
