"""Turn ncu outputs under gpurun_out/ into the tracked summaries under profiles/.

  python tools/summarize_ncu.py launches gpurun_out/launches_n20k.csv profiles/r01_launches_n20k.md "<command>"
  python tools/summarize_ncu.py full gpurun_out/prof_x.ncu-rep profiles/r01_syrk_ws_n20k.md
"""
import collections
import csv
import io
import subprocess
import sys


def launches(src, dst, cmd):
    lines = [l for l in open(src) if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    agg = collections.OrderedDict()
    half = len(rows) // 2           # profile_eval.py runs warm-up + 1 evaluation: keep the second
    for r in rows[half:]:
        name = r["Kernel Name"].split("(")[0].replace("<unnamed>::", "")
        t = float(r["Metric Value"].replace(",", ""))
        a = agg.setdefault(name, [0, 0.0, 0.0])
        a[0] += 1; a[1] += t; a[2] = max(a[2], t)
    tot = sum(v[1] for v in agg.values())
    with open(dst, "w") as f:
        f.write("# ncu launch list (gpu__time_duration.sum, --clock-control none)\n\n")
        f.write("command: `%s`\n\nSecond evaluation of the run (%d launches). Times are cold-cache and serialised by ncu: "
                "compare SHARES, not absolutes.\n\n" % (cmd, len(rows) - half))
        f.write("| kernel | launches | total ms | share | avg us | max us |\n|---|---:|---:|---:|---:|---:|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("| %s | %d | %.3f | %.1f%% | %.1f | %.1f |\n" % (k, v[0], v[1] / 1e6, 100 * v[1] / tot, v[1] / v[0] / 1e3, v[2] / 1e3))
        f.write("| **total** | %d | %.3f | 100%% | | |\n" % (len(rows) - half, tot / 1e6))
    print(open(dst).read())


KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.max",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tma_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed"]


def full(src, dst):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, body = rows[0], rows[1], rows[2:]
    ci = {h: i for i, h in enumerate(hdr)}
    with open(dst, "w") as f:
        f.write("# ncu --set full summary: %s\n\n" % src)
        for r in body:
            f.write("## %s  grid %s block %s\n\n| metric | value | unit |\n|---|---:|---|\n" % (
                r[ci["Kernel Name"]], r[ci["Grid Size"]], r[ci["Block Size"]]))
            for k in KEYS:
                if k in ci:
                    f.write("| %s | %s | %s |\n" % (k, r[ci[k]], units[ci[k]]))
            stalls = [(h, float(r[i])) for h, i in ci.items() if h.startswith("smsp__pcsamp_warps_issue_stalled_")
                      and not h.endswith("_not_issued") and r[i] not in ("", "n/a")]
            tot = sum(v for _, v in stalls) or 1.0
            f.write("\nwarp-state samples (share): " + ", ".join(
                "%s %.1f%%" % (h.replace("smsp__pcsamp_warps_issue_stalled_", ""), 100 * v / tot)
                for h, v in sorted(stalls, key=lambda kv: -kv[1])[:8]) + "\n\n")
    print(open(dst).read())


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else "")
    else:
        full(sys.argv[2], sys.argv[3])
