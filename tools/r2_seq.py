"""Sequence of launches (name, us) from an ncu duration list: the last `count` launches."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1], errors="replace")))
count = int(sys.argv[2]) if len(sys.argv) > 2 else 80
h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr = rows[h]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
gi = hdr.index("Grid Size") if "Grid Size" in hdr else None
seq = [(r[ki], float(r[vi].replace(",", "")), r[gi] if gi is not None else "") for r in rows[h + 1:] if len(r) > vi]
for name, v, grid in seq[-count:]:
    short = name.replace("cugp::<unnamed>::", "").replace("void ", "").split("(")[0][:60]
    print(f"{v/1e3:9.1f} us  {grid:>16}  {short}")
print(f"total {sum(v for _, v, _ in seq[-count:])/1e3:.1f} us over {min(count, len(seq))} launches")
