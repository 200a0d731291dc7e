"""Summarise ncu captures into a markdown file under profiles/.

    python tools/ncu_summary.py <launches.csv> <full.ncu-rep> <out.md> [title]
"""
import collections, csv, io, os, subprocess, sys

launch_csv, rep, out = sys.argv[1:4]
title = sys.argv[4] if len(sys.argv) > 4 else os.path.basename(out)

# ---- launch list
rows = [r for r in csv.reader(open(launch_csv)) if len(r) > 10]
hdr = None
per = collections.OrderedDict()
for r in rows:
    if "Kernel Name" in r:
        hdr = r
        continue
    if hdr is None or len(r) != len(hdr):
        continue
    name = r[hdr.index("Kernel Name")]
    if not name.startswith(("kws::", "void kws::")):
        continue
    name = name.replace("void ", "").split("(")[0]
    ns = float(r[hdr.index("Metric Value")].replace(",", ""))
    per.setdefault(name, []).append(ns)

def steady(v):                                      # drop warm-up outliers: median of the last half
    v = sorted(v[len(v) // 2:])
    return v[len(v) // 2]

L = []
L.append("# %s\n" % title)
L.append("## Launch list (`ncu --metrics gpu__time_duration.sum --clock-control none`; cold-cache, serialised: compare shares)\n")
L.append("| kernel | launches | median ms (2nd half) | share of one step |")
L.append("|---|---|---|---|")
tot = sum(steady(v) for v in per.values())
for k, v in per.items():
    L.append("| `%s` | %d | %.3f | %.1f%% |" % (k, len(v), steady(v) / 1e6, 100 * steady(v) / tot))
L.append("\nSum of one launch of each = %.3f ms.\n" % (tot / 1e6))

# ---- full capture
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
rr = list(csv.reader(io.StringIO(raw)))
h, units, data = rr[0], rr[1], rr[2:]
want = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
]
names = [r[h.index("Kernel Name")].split("(")[0].replace("void ", "") for r in data]
L.append("## `ncu --set full --clock-control none --import-source on` (one launch each; %s)\n" % os.path.basename(rep))
L.append("| metric | " + " | ".join("`%s`" % n for n in names) + " |")
L.append("|---|" + "---|" * len(names))
for w in want:
    if w not in h:
        continue
    i = h.index(w)
    L.append("| %s (%s) | " % (w, units[i]) + " | ".join(r[i] for r in data) + " |")
open(out, "w").write("\n".join(L) + "\n")
print("\n".join(L))
