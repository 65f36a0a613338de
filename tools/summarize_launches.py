"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1], errors="replace")))
h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr = rows[h]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
tot, cnt = collections.Counter(), collections.Counter()
for r in rows[h + 1:]:
    if len(r) <= vi:
        continue
    name = r[ki].replace("cugp::<unnamed>::", "").replace("void ", "").split("(cugp")[0].split("(const")[0][:80]
    tot[name] += float(r[vi].replace(",", ""))
    cnt[name] += 1
total = sum(tot.values())
print(f"{'us':>12} {'share':>7} {'count':>6}  kernel")
for k, v in tot.most_common():
    print(f"{v/1e3:12.1f} {100*v/total:6.2f}% {cnt[k]:6d}  {k}")
print(f"{total/1e3:12.1f} us total, {sum(cnt.values())} launches")
