"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list (last iteration's worth of
launches when the list holds several identical iterations is NOT separated: totals are over the whole file)."""
import csv
import collections
import re
import sys

rows = list(csv.reader(l for l in open(sys.argv[1]) if not l.startswith("==")))
hdr = rows[0]
ki, mi, vi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
tot, cnt = collections.Counter(), collections.Counter()
for r in rows[1:]:
    if len(r) <= vi or r[mi] != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r[ki])
    name = re.sub(r"^void ", "", name)
    tot[name] += float(r[vi].replace(",", ""))
    cnt[name] += 1
total = sum(tot.values())
print("total %.3f ms over %d launches" % (total / 1e6, sum(cnt.values())))
for n, t in tot.most_common(40):
    print("%8.3f ms %5.1f%% %6d  %s" % (t / 1e6, 100 * t / total, cnt[n], n[:110]))
