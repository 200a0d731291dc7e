"""Hot spots of one kernel of an ncu report (SASS view): instructions ranked by warp-stall samples, with the dominant
stall reasons, plus totals per opcode.  usage: python tools/ncu_hot.py SOURCE_PAGE.csv KERNEL_SUBSTRING [top]"""
import collections, csv, sys

rows = list(csv.reader(open(sys.argv[1])))
kern = sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
start = next(i for i, r in enumerate(rows) if r and r[0] == "Kernel Name" and kern in r[1])
hdr = rows[start + 1]
data = []
for r in rows[start + 2:]:
    if r and r[0] == "Kernel Name":
        break
    if len(r) == len(hdr):
        data.append(r)
i_s, i_e, i_src = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Source")
stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h and "not_issued" not in h.lower()]
tot_s = sum(int(r[i_s]) for r in data)
tot_e = sum(int(r[i_e]) for r in data)
print("instructions %d, executed %d, samples %d" % (len(data), tot_e, tot_s))
reasons = collections.Counter()
for r in data:
    for i in stall:
        reasons[hdr[i]] += int(r[i])
print("stall reasons:", [(k, v) for k, v in reasons.most_common(12)])
ranked = sorted(range(len(data)), key=lambda k: -int(data[k][i_s]))[:top]
for k in sorted(ranked):
    r = data[k]
    rs = sorted(((int(r[i]), hdr[i][6:]) for i in stall), reverse=True)[:2]
    print("%5d %6s smp %8s exe  %-70s %s" % (k, r[i_s], r[i_e], r[i_src].strip()[:70], rs))
