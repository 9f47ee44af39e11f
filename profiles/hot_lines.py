"""Per-CUDA-source-line hot spots from an .ncu-rep captured with --import-source on (-lineinfo build).
    python profiles/hot_lines.py gpurun_out/prof.ncu-rep [top_n]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                     stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur_file, hdr, lines = None, None, []
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
    elif len(r) > 5 and r[0] == "Line No":
        hdr = {k: i for i, k in enumerate(r)}
    elif hdr and len(r) > 10 and r[2] == "-":
        def g(k):
            try: return int(r[hdr[k]])
            except Exception: return 0
        stalls = {k[6:]: g(k) for k in hdr if k.startswith("stall_") and "Not Issued" not in k}
        lines.append((g("# Samples"), g("Instructions Executed"), cur_file, r[0], r[1].strip()[:90], stalls))
tot = sum(l[0] for l in lines) or 1
toti = sum(l[1] for l in lines) or 1
print("total samples %d, total warp-instructions %d" % (tot, toti))
for s, ie, f, ln, src, st in sorted(lines, key=lambda l: -l[0])[:top]:
    top3 = ", ".join("%s %d%%" % (k, 100 * v // max(s, 1)) for k, v in sorted(st.items(), key=lambda x: -x[1])[:3])
    print("%5.1f%% smp %5.1f%% inst  %-14s:%-4s %-90s [%s]" % (100.0 * s / tot, 100.0 * ie / toti, f, ln, src, top3))

# ---- stall mix of the gating / set-up arithmetic (ekf_small.cuh), i.e. of the warps others wait for ---
import collections
agg = collections.defaultdict(collections.Counter)
inst = collections.Counter()
for s_, ie, f, ln, src, st in lines:
    if f == "ekf_small.cuh":
        r = "small:gating(85-225)" if 85 <= int(ln) <= 225 else "small:prop/setup"
        agg[r].update(st)
        inst[r] += ie
for r, c in agg.items():
    t = sum(c.values()) or 1
    print("%-22s samples %6d inst %.2e :" % (r, t, inst[r]), ", ".join("%s %d%%" % (k, 100 * v // t) for k, v in c.most_common(6)))
