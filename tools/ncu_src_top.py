"""Top stall instructions and opcode histogram of an `ncu --page source --csv` dump: python tools/ncu_src_top.py dump.csv [n]"""
import csv, sys
from collections import Counter
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = next(r for r in rows if r and r[0] == "Address")
ix = {h: i for i, h in enumerate(hdr)}
def num(r, k):
    try: return int(r[ix[k]])
    except Exception: return 0
data = [r for r in rows if len(r) > 5 and r[0].startswith("0x")]
tot = sum(num(r, "# Samples") for r in data)
print("total samples", tot, "instructions", len(data))
keys = [k for k in hdr if k.startswith("stall_") and "Not Issued" not in k]
for r in sorted(data, key=lambda r: -num(r, "# Samples"))[:n]:
    big = sorted(((k, num(r, k)) for k in keys), key=lambda kv: -kv[1])[:2]
    print(str(num(r, "# Samples")).rjust(7), str(num(r, "Instructions Executed")).rjust(9), r[ix["Source"]].strip()[:72].ljust(72), big)
c, s = Counter(), Counter()
for r in data:
    parts = r[ix["Source"]].strip().split()
    op = parts[1] if parts[0].startswith("@") else parts[0]
    c[op.split(".")[0]] += num(r, "Instructions Executed"); s[op.split(".")[0]] += num(r, "# Samples")
print([(k, v, s[k]) for k, v in c.most_common(25)])
