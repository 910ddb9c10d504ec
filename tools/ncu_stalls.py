"""dev helper: per-region stall-reason breakdown from an ncu source-page CSV (tools/ncu_stalls.py src.csv)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
H = rows[1]
ix = {h: i for i, h in enumerate(H)}
stall_cols = [h for h in H if h.startswith("stall_") and "Not Issued" not in h]
body = [r for r in rows[2:] if len(r) == len(H)]
def iv(r, h):
    try: return int(r[ix[h]])
    except: return 0
tot = sum(iv(r, "# Samples") for r in body)
totals = {h: sum(iv(r, h) for r in body) for h in stall_cols}
print("total samples", tot, {h[6:]: v for h, v in totals.items() if v > 0.01 * tot})
print("inst executed", sum(iv(r, "Instructions Executed") for r in body))
# regions split at marker instructions
marks = ('BAR.SYNC', 'SYNCS.ARRIVE', 'EXIT', 'FENCE.VIEW.ASYNC', 'UBLKCP')
start = 0; acc = {h: 0 for h in stall_cols}; ns = 0; ni = 0
for i, r in enumerate(body):
    ns += iv(r, "# Samples"); ni += iv(r, "Instructions Executed")
    for h in stall_cols: acc[h] += iv(r, h)
    if any(m in r[ix["Source"]] for m in marks) and (ns > 0.004 * tot or 'EXIT' in r[ix["Source"]]):
        top = sorted(acc.items(), key=lambda kv: -kv[1])[:5]
        print(f"[{start:4d}-{i:4d}] samples {ns:6d} ({100*ns/tot:4.1f}%) inst {ni:9d} | " +
              " ".join(f"{h[6:]}={v}" for h, v in top if v) + " | " + r[ix["Source"]].strip()[:40])
        start = i + 1; acc = {h: 0 for h in stall_cols}; ns = 0; ni = 0
if len(sys.argv) > 2:
    a, b = int(sys.argv[2]), int(sys.argv[3])
    for i in range(a, b):
        r = body[i]
        top = sorted(((h, iv(r, h)) for h in stall_cols), key=lambda kv: -kv[1])[:3]
        print(f"{i:4d} {iv(r,'# Samples'):5d} {iv(r,'Instructions Executed'):9d} " + " ".join(f"{h[6:]}={v}" for h, v in top if v).ljust(40) + " " + r[ix["Source"]].strip()[:90])
