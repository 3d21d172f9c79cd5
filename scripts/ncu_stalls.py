"""Top stall sites of one kernel in an .ncu-rep (source page, SASS view).  Usage: ncu_stalls.py rep kernel [topN]"""
import csv, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
top_n = int(sys.argv[3]) if len(sys.argv) > 3 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", kern, "--launch-count", "1"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
h = rr[1]; rows = []
for r in rr[2:]:
    if r and r[0] in ("Kernel Name", "Address"): break
    if len(r) == len(h): rows.append(r)
iS = h.index("# Samples"); isrc = h.index("Source"); iex = h.index("Instructions Executed")
cols = [c for c in h if c.startswith("stall_") and "Not Issued" not in c]
tot = sum(int(r[iS]) for r in rows)
print("instr", len(rows), "samples", tot)
for c in cols:
    v = sum(int(r[h.index(c)]) for r in rows)
    if v * 50 > tot: print("  %-22s %6d  %.1f%%" % (c, v, 100.0 * v / tot))
top = sorted(range(len(rows)), key=lambda i: -int(rows[i][iS]))[:top_n]
for i in sorted(top):
    r = rows[i]
    why = max(cols, key=lambda c: int(r[h.index(c)]))
    print("%5d %6s %-14s x%-8s %s" % (i, r[iS], why[6:], r[iex], r[isrc][:90]))
