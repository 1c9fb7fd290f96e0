"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import collections
import csv
import sys

src, dst = sys.argv[1], sys.argv[2]
with open(src) as f:
    lines = [l for l in f if not l.startswith("==")]
tot = collections.OrderedDict()
for r in csv.DictReader(lines):
    try:
        v = float(r["Metric Value"].replace(",", ""))
    except Exception:
        continue
    u = r["Metric Unit"]
    v = v / 1e3 if u in ("ns", "nsecond") else (v if u in ("us", "usecond") else v * 1e3 if u in ("ms", "msecond") else v)
    t = tot.setdefault(r["Kernel Name"][:90], [0, 0.0])
    t[0] += 1
    t[1] += v
s = sum(v[1] for v in tot.values())
out = [f"total {s / 1e3:.2f} ms over {sum(v[0] for v in tot.values())} launches (setup + warm-up + timed steps, serialised and cold-cache under ncu)"]
for k, (n, v) in sorted(tot.items(), key=lambda kv: -kv[1][1])[:45]:
    out.append(f"{v / 1e3:10.3f} ms {100 * v / s:5.1f}%  x{n:5d}  {k}")
open(dst, "w").write("\n".join(out) + "\n")
print("\n".join(out))
