"""Summarise an ncu launch list (`ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file X`) as a
markdown table: total time, share and launch count per kernel.   python tools/launch_summary.py X.csv "title" > out.md"""
import csv, sys
from collections import defaultdict

path, title = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "ncu launch list")
rows = [l for l in open(path, newline="") if l.startswith('"')]
rd = csv.DictReader(rows)
tot, cnt = defaultdict(float), defaultdict(int)
for r in rd:
    if r["Metric Name"] != "gpu__time_duration.sum":
        continue
    scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0}.get(r["Metric Unit"], 1e-6)
    k = r["Kernel Name"]
    tot[k] += float(r["Metric Value"].replace(",", "")) * scale
    cnt[k] += 1
total = sum(tot.values())
print(f"# {title}\n")
print("`ncu --metrics gpu__time_duration.sum --clock-control none --csv` after the same command exited 0 without ncu")
print("(cold-cache, serialised: compare shares, not absolutes).")
print(f"{sum(cnt.values())} launches, {total:.3f} ms total.\n")
print("| ms | share | launches | kernel |\n|---|---|---|---|")
for k in sorted(tot, key=tot.get, reverse=True):
    if tot[k] / total < 0.002:
        continue
    print(f"| {tot[k]:.3f} | {100 * tot[k] / total:.1f}% | {cnt[k]} | `{k[:90]}` |")
