"""Turn ncu outputs (gpurun_out/*.ncu-rep, launches*.csv) into the small text summaries committed
under profiles/. Runs here (no GPU needed):

    python profiles/summarize.py rep  gpurun_out/prof_x.ncu-rep  profiles/r01_x          # -> _metrics.csv, _hot.txt
    python profiles/summarize.py list gpurun_out/launches.csv    profiles/r01_launches.md
"""
import csv
import io
import subprocess
import sys
from collections import defaultdict

KEEP = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
    "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
    "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "smsp__average_warp_latency_per_inst_issued.ratio",
]
KEEP_PREFIX = ["smsp__average_warps_issue_stalled_"]


def ncu_csv(rep, page, extra=()):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv", *extra], stdout=subprocess.PIPE,
                         stderr=subprocess.DEVNULL, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def summarize_rep(rep, prefix):
    rows = ncu_csv(rep, "raw")
    hdr, units = rows[0], rows[1]
    with open(prefix + "_metrics.csv", "w") as f:
        w = csv.writer(f)
        w.writerow(["kernel", "metric", "unit", "value"])
        for vals in rows[2:]:
            name = vals[hdr.index("Kernel Name")]
            for h, u, v in zip(hdr, units, vals):
                if h in KEEP or any(h.startswith(p) for p in KEEP_PREFIX):
                    w.writerow([name, h, u, v])
    # SASS hot spots of the first kernel in the report
    src = ncu_csv(rep, "source")
    hi = next(i for i, r in enumerate(src) if len(r) > 4 and r[0] == "Address")
    h = src[hi]
    ix = {k: i for i, k in enumerate(h)}
    data = [r for r in src[hi + 1:] if len(r) == len(h) and r[ix["# Samples"]].isdigit()]
    tot = sum(int(r[ix["# Samples"]] or 0) for r in data) or 1
    stall_cols = [c for c in h if c.startswith("stall_") and "Not Issued" not in c]
    with open(prefix + "_hot.txt", "w") as f:
        f.write("kernel: %s\n" % src[0][1] if src and len(src[0]) > 1 else "")
        f.write("total warp-stall samples: %d over %d SASS instructions\n\n" % (tot, len(data)))
        f.write("per 50-instruction window: share of samples, executed warp-instructions, dominant opcodes, dominant stalls\n")
        for c in range(0, len(data), 50):
            seg = data[c:c + 50]
            s = sum(int(r[ix["# Samples"]] or 0) for r in seg)
            ie = sum(int(r[ix["Instructions Executed"]] or 0) for r in seg)
            ops = defaultdict(int)
            st = defaultdict(int)
            for r in seg:
                t = r[ix["Source"]].split()
                if t:
                    ops[t[1] if t[0].startswith("@") and len(t) > 1 else t[0]] += 1
                for sc in stall_cols:
                    st[sc] += int(r[ix[sc]] or 0)
            if s * 200 < tot:
                continue
            f.write("%5d  %5.1f%%  inst=%.2e  ops=%s  stalls=%s\n" % (
                c, 100.0 * s / tot, ie, dict(sorted(ops.items(), key=lambda x: -x[1])[:5]),
                dict(sorted(st.items(), key=lambda x: -x[1])[:3])))
        f.write("\ntop 25 single instructions by samples\n")
        for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]] or 0))[:25]:
            f.write("%6.2f%%  %s\n" % (100.0 * int(r[ix["# Samples"]] or 0) / tot, r[ix["Source"]]))


def summarize_list(path, out):
    rows = list(csv.reader(open(path, errors="replace")))
    hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    h = rows[hi]
    kn, mv, mn = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Name")
    agg = defaultdict(lambda: [0, 0.0])
    unit = ""
    for r in rows[hi + 1:]:
        if len(r) != len(h) or r[mn] != "gpu__time_duration.sum":
            continue
        unit = r[h.index("Metric Unit")]
        a = agg[r[kn].split("(")[0]]
        a[0] += 1
        a[1] += float(r[mv].replace(",", ""))
    tot = sum(a[1] for a in agg.values()) or 1.0
    with open(out, "w") as f:
        f.write("ncu launch list (%s): gpu__time_duration.sum per kernel, cold-cache and serialised - compare SHARES\n\n" % path)
        f.write("| kernel | launches | total (%s) | share | avg (%s) |\n|---|---:|---:|---:|---:|\n" % (unit, unit))
        for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write("| %s | %d | %.1f | %.1f%% | %.2f |\n" % (k, n, t, 100 * t / tot, t / n))


if __name__ == "__main__":
    mode, src, dst = sys.argv[1:4]
    (summarize_rep if mode == "rep" else summarize_list)(src, dst)
