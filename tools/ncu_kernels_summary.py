"""dev helper: ncu CSV (gpu__time_duration, dram bytes per launch) -> per-kernel table:
last launch of every kernel name, achieved DRAM GB/s and fraction of the measured HBM peak."""
import csv, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
peak = 6540.5
p = os.path.join(ROOT, "MEASURED_PEAKS.json")
if os.path.exists(p):
    peak = float(json.load(open(p))["hbm_gbs"])
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 6 and r[0].isdigit()]
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0, "usecond": 1e-6, "nsecond": 1e-9, "msecond": 1e-3, "second": 1.0}
launches = {}
for r in rows:
    lid, name, metric, unit, val = r[0], r[4], r[-3], r[-2], float(r[-1].replace(",", ""))
    launches.setdefault(lid, {"name": name})[metric] = val * UNIT.get(unit, 1)
last = {}
for lid in sorted(launches, key=int):
    d = launches[lid]
    key = d["name"].split("(")[0]
    last[key] = d
    last[key]["count"] = last.get(key, {}).get("count", 0)
counts = {}
for d in launches.values():
    k = d["name"].split("(")[0]
    counts[k] = counts.get(k, 0) + 1
print(f"# per-kernel ncu evidence (last launch of each kernel; cold-cache, serialised) — HBM peak {peak} GB/s (measured)")
print(f"{'kernel':70s} {'launches':>8s} {'us':>9s} {'read MB':>9s} {'write MB':>9s} {'DRAM GB/s':>10s} {'of peak':>8s}")
for k, d in sorted(last.items(), key=lambda kv: -kv[1].get("gpu__time_duration.sum", 0)):
    t = d.get("gpu__time_duration.sum", 0.0)
    rd, wr = d.get("dram__bytes_read.sum", 0.0), d.get("dram__bytes_write.sum", 0.0)
    gbs = (rd + wr) / t / 1e9 if t > 0 else 0.0
    print(f"{k[:70]:70s} {counts[k]:8d} {t * 1e6:9.1f} {rd / 1e6:9.1f} {wr / 1e6:9.1f} {gbs:10.1f} {gbs / peak:8.3f}")
