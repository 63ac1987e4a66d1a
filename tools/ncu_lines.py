#!/usr/bin/env python
"""Aggregate `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass` per CUDA source line.
usage: ncu_lines.py file.csv [top_n]   -- prints, per file, the lines with most stall samples / instructions."""
import csv, sys, collections
path = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(open(path)))
fname = None; hdr = None; cur = None
agg = collections.OrderedDict()
for r in rows:
    if not r: continue
    if r[0] == "File Name": fname = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; continue
    if r[0] in ("Kernel Name", "File Path") or (r[0] != "" and not r[0].isdigit()): continue
    if hdr is None: continue
    d = dict(zip(range(len(hdr)), r))
    if r[0] != "":            # a CUDA source line
        cur = (fname, int(r[0])); agg.setdefault(cur, dict(src=r[1], samples=0, inst=0, stalls=collections.Counter()))
        continue
    if cur is None: continue
    a = agg[cur]
    def num(x):
        try: return float(x)
        except Exception: return 0.0
    a["samples"] += num(r[hdr.index("# Samples")]); a["inst"] += num(r[hdr.index("Instructions Executed")])
    for i, h in enumerate(hdr):
        if h.startswith("stall_") and "Not Issued" not in h:
            a["stalls"][h[6:]] += num(r[i])
tot_s = sum(a["samples"] for a in agg.values()); tot_i = sum(a["inst"] for a in agg.values())
print("total samples %d, warp instructions %d" % (tot_s, tot_i))
st = collections.Counter()
for a in agg.values(): st.update(a["stalls"])
print("stalls:", ", ".join("%s %.1f%%" % (k, 100 * v / max(1, sum(st.values()))) for k, v in st.most_common(10)))
print("---- by samples")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[:top]:
    s2 = ",".join("%s:%d" % (n, v) for n, v in a["stalls"].most_common(3))
    print("%-18s %5d  %5.1f%% smp %5.1f%% inst  %-40s | %s" % (k[0], k[1], 100 * a["samples"] / tot_s, 100 * a["inst"] / tot_i, s2, a["src"].strip()[:90]))
