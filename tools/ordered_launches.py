"""Print the launches of ONE optimiser step from an ncu launch list in order, with durations (us):
python tools/ordered_launches.py launches.csv [anchor-kernel-substring]  -- the step is the span between the last two
launches of the anchor (default: gather_rows, the first kernel of fit())."""
import csv, sys
path = sys.argv[1]
anchor = sys.argv[2] if len(sys.argv) > 2 else "gather_rows"
rows = [l for l in open(path, newline="") if l.startswith('"')]
out = []
for r in csv.DictReader(rows):
    if r["Metric Name"] != "gpu__time_duration.sum":
        continue
    scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r["Metric Unit"], 1e-3)
    out.append((r["Kernel Name"], float(r["Metric Value"].replace(",", "")) * scale))
idx = [i for i, (k, _) in enumerate(out) if anchor in k]
# gather_rows is launched twice per fit (relations, subjects): take the first of each pair
starts = [i for n, i in enumerate(idx) if n == 0 or i - idx[n - 1] > 3]
a, b = starts[-2], starts[-1]
tot = 0.0
for k, t in out[a:b]:
    tot += t
    print(f"{t:9.1f}  {k[:110]}")
print(f"--- {b - a} launches, {tot:.1f} us")
