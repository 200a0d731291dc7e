"""Per-phase (barrier to barrier) instruction and stall-sample breakdown of one kernel of an ncu report.

usage: python tools/ncu_phases.py REPORT.ncu-rep KERNEL_REGEX [items]
Reads `ncu --page source --csv` (SASS view) and cuts the instruction stream at every BAR.SYNC.
"""
import collections
import csv
import io
import subprocess
import sys


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    items = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern],
                         stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr, data = rows[hdr_i], []
    for r in rows[hdr_i + 1:]:
        if r and r[0] == "Kernel Name":
            break                                   # first matching launch only
        if len(r) == len(hdr):
            data.append(r)
    i_s, i_e, i_src = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Source")
    stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    phase = 0
    agg = collections.defaultdict(lambda: [0, 0, collections.Counter(), collections.Counter()])
    for r in data:
        a = agg[phase]
        a[0] += int(r[i_s])
        a[1] += int(r[i_e])
        ins = r[i_src].split()
        op = ins[1] if ins and ins[0].startswith("@") and len(ins) > 1 else (ins[0] if ins else "?")
        a[2][op] += int(r[i_e])
        for i in stall:
            a[3][hdr[i][6:]] += int(r[i])
        if "BAR.SYNC" in r[i_src]:
            phase += 1
    ts = sum(v[0] for v in agg.values()) or 1
    te = sum(v[1] for v in agg.values()) or 1
    print("total warp-instructions %d (%.0f per item), samples %d" % (te, te / items, ts))
    for k in sorted(agg):
        v = agg[k]
        print("phase %d: samples %4.1f%%  inst %4.1f%% (%.0f/item)  stalls: %s\n          ops: %s" % (
            k, 100.0 * v[0] / ts, 100.0 * v[1] / te, v[1] / items,
            ", ".join("%s %d" % kv for kv in v[3].most_common(4)),
            ", ".join("%s %.0f" % (o, c / items) for o, c in v[2].most_common(9))))


if __name__ == "__main__":
    main()
