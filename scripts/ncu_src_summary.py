"""Summarise `ncu --page source --csv` output: per-kernel opcode mix, stall samples, hot instructions."""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
kernels = []
cur = None
for r in rows:
    if len(r) >= 2 and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "data": []}
        kernels.append(cur)
    elif cur is not None and cur["hdr"] is None:
        cur["hdr"] = r
    elif cur is not None and len(r) == len(cur["hdr"]):
        cur["data"].append(r)
for ki, k in enumerate(kernels):
    hdr, data = k["hdr"], k["data"]
    iS, iSamp, iEx = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    tot_ex = sum(int(r[iEx]) for r in data)
    tot_s = sum(int(r[iSamp]) for r in data)
    print("== kernel %d: %d static instr, %d executed static, warp-instr %d, samples %d" % (
        ki, len(data), sum(1 for r in data if int(r[iEx]) > 0), tot_ex, tot_s))
    op, ops = collections.Counter(), collections.Counter()
    for r in data:
        t = r[iS].split()
        o = t[1] if t[0].startswith("@") else t[0]
        o = o.split(".")[0]
        op[o] += int(r[iEx]); ops[o] += int(r[iSamp])
    print("  " + " | ".join("%s %.1f%%/%.1f%%" % (o, 100 * c / tot_ex, 100 * ops[o] / max(tot_s, 1)) for o, c in op.most_common(22)))
    stall = collections.Counter()
    for i, h in enumerate(hdr):
        if h.startswith("stall_"):
            for r in data:
                try: stall[h] += int(r[i])
                except ValueError: pass
    print("  stalls: " + ", ".join("%s %d" % (h, c) for h, c in stall.most_common(10)))
    if len(sys.argv) > 2:
        top = sorted(data, key=lambda r: -int(r[iSamp]))[:int(sys.argv[2])]
        for r in top:
            print("   %6s samples  exec %9s  %s" % (r[iSamp], r[iEx], r[iS]))
