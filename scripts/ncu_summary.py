"""Summarise ncu outputs brought back in gpurun_out/ into profiles/ (tracked).
usage: python scripts/ncu_summary.py <launches.csv> <prof.ncu-rep|-> <tag>"""
import collections, csv, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
launch_csv, rep, tag = sys.argv[1], sys.argv[2], sys.argv[3]
out = []
rows = list(csv.reader(open(launch_csv)))
hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr = rows[hi]; ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) <= vi: continue
    a = agg.setdefault(r[ki].split("(")[0], [0, 0.0]); a[0] += 1; a[1] += float(r[vi].replace(",", ""))
tot = sum(a[1] for a in agg.values())
out.append("# ncu launch list (%s): gpu__time_duration.sum per kernel, --clock-control none" % tag)
out.append("# cold-cache, serialised: compare SHARES, not absolutes")
out.append("%-44s %5s %12s %7s %10s" % ("kernel", "n", "total_us", "share", "avg_us"))
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append("%-44s %5d %12.1f %6.1f%% %10.1f" % (k[:44], n, t / 1e3, 100 * t / tot, t / n / 1e3))
out.append("total_us %.1f" % (tot / 1e3))
if rep != "-":
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(raw.splitlines()))
    h = rr[0]
    want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_tensor.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__grid_size",
            "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_bytes.sum"]
    out.append("")
    out.append("# ncu --set full, %s (units in row 2)" % os.path.basename(rep))
    for w in want:
        if w in h:
            i = h.index(w)
            out.append("%-66s %s" % (w, " | ".join(r[i] for r in rr[1:5])))
path = os.path.join(ROOT, "profiles", "ncu_%s.txt" % tag)
open(path, "w").write("\n".join(out) + "\n")
print("\n".join(out))
