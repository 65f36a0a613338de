"""profiles/top_kernel_traffic.json from an ncu raw page of the Cholesky's trailing-update launches: the longest
captured launch (the first full-width K = outer-width update of tools/chol_only.py) with its DRAM traffic next to
its algorithmic bytes.  usage: make_traffic_json.py profiles/r1_ncu_trailing_raw.csv n outer_width"""
import csv
import json
import sys

path, n, nb = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
rows = list(csv.reader(open(path, errors="replace")))
hdr, units = rows[0], rows[1]


def val(d, k, scale_to=None):
    unit, v = d[k]
    v = float(v.replace(",", ""))
    mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}
    return v * mult.get(unit, 1.0)


best = None
for vals in rows[2:]:
    d = dict(zip(hdr, zip(units, vals)))
    t = val(d, "gpu__time_duration.sum")
    if best is None or t > best[0]:
        best = (t, d)
t, d = best
m = n + 1 - 2 * nb          # rows of the first U2 update (incl. the appended y^T row): right of panels 0 and 1
w = n - 2 * nb
tiles = (w // 128) * (w // 128 + 1) // 2 + ((m + 127) // 128 - w // 128) * (w // 128)
alg = 2 * 8 * (w * (w + 1) // 2 + (m - w) * w) + 8 * m * nb
out = {
    "capture": f"{path} (ncu --set full --clock-control none, tools/chol_only.py {n}: longest captured launch = first "
               f"full-width K={nb} trailing update U2(0), m = {m} rows x {w} columns, lower tiles)",
    "grid": int(float(d["launch__grid_size"][1])),
    "duration_ms": t,
    "dram_bytes_read": val(d, "dram__bytes_read.sum"),
    "dram_bytes_write": val(d, "dram__bytes_write.sum"),
    "dram_bytes_total": val(d, "dram__bytes_read.sum") + val(d, "dram__bytes_write.sum"),
    "algorithmic_bytes": alg,
    "note": "algorithmic = read + write of every lower C element once + the m x K panel once",
    "tensor_pipe_active_pct": float(d["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"][1]),
    "l2_hit_rate_pct": float(d["lts__t_sector_hit_rate.pct"][1]),
}
json.dump(out, open("profiles/top_kernel_traffic.json", "w"), indent=1)
print(json.dumps(out, indent=1))
