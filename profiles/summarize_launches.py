"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel: python profiles/summarize_launches.py csv"""
import csv
import re
import sys
from collections import OrderedDict

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if not l.startswith("=="))]
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = OrderedDict()
for r in rows[1:]:
    if len(r) <= vi or not r[vi]:
        continue
    name = re.sub(r"\(.*", "", r[ki]).strip()
    v = float(r[vi].replace(",", ""))
    us = v / 1e3 if r[ui] in ("ns", "nsecond") else (v if r[ui] in ("us", "usecond") else v * 1e3)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += us
ours = {k: v for k, v in agg.items() if "ieee::" in k}
tot = sum(v[1] for v in ours.values())
print("%-62s %8s %12s %7s %10s" % ("kernel", "launches", "total us", "share", "avg us"))
for k, (n, us) in sorted(ours.items(), key=lambda kv: -kv[1][1]):
    print("%-62s %8d %12.1f %6.1f%% %10.1f" % (k[:62], n, us, 100 * us / tot, us / n))
other = {k: v for k, v in agg.items() if "ieee::" not in k}
print("(not this library's: %s)" % ", ".join("%s x%d %.0f us" % (k[:40], n, us) for k, (n, us) in other.items()))
