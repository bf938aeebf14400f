"""Summarise an ncu --csv launch list (gpu__time_duration.sum) by kernel name."""
import csv, sys, re, collections
rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
rd = csv.reader(lines)
hdr = next(rd)
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rd:
    if len(r) <= vi: continue
    name = re.sub(r"\(.*", "", r[ki]); name = re.sub(r"^void ", "", name)
    v = float(r[vi].replace(",", "")); u = r[ui]
    us = v / 1e3 if u in ("ns", "nsecond") else (v if u in ("us", "usecond") else v * 1e3)
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += us
tot = sum(a[1] for a in agg.values())
print("total %.2f ms over %d launches" % (tot / 1e3, sum(a[0] for a in agg.values())))
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print("%9.2f ms %5.1f%% %6d x %8.1f us  %s" % (a[1] / 1e3, 100 * a[1] / tot, a[0], a[1] / a[0], k[:110]))
