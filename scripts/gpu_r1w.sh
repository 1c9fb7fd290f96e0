#!/bin/bash
# where does the 50 ms training step go: GEMM microbench at the config-4 shapes + per-kernel launch list of the train bench
mkdir -p gpurun_out
timeout 300 python scripts/gemm_bench.py > gpurun_out/r1w_gemm_bench.jsonl 2> gpurun_out/r1w_gemm.err; echo "gemm exit=$?"; cat gpurun_out/r1w_gemm_bench.jsonl | cut -c1-200
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r1w_train_launches.csv python bench.py --workload train --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r1w_train_under_ncu.log 2>&1; echo "ncu exit=$?"
python - <<'PY'
import csv, collections
rows = []
with open("gpurun_out/r1w_train_launches.csv") as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
tot = collections.OrderedDict()
for r in rd:
    try:
        v = float(r["Metric Value"].replace(",", ""))
    except Exception:
        continue
    u = r["Metric Unit"]
    v = v / 1e3 if u in ("ns", "nsecond") else (v if u in ("us", "usecond") else v * 1e3 if u in ("ms", "msecond") else v)
    k = r["Kernel Name"][:70]
    t = tot.setdefault(k, [0, 0.0]); t[0] += 1; t[1] += v
s = sum(v[1] for v in tot.values())
out = [f"total {s/1e3:.2f} ms over {sum(v[0] for v in tot.values())} launches (warmup + timed steps + setup, serialised under ncu)"]
for k, (n, v) in sorted(tot.items(), key=lambda kv: -kv[1][1])[:40]:
    out.append(f"{v/1e3:9.3f} ms {100*v/s:5.1f}%  x{n:5d}  {k}")
open("gpurun_out/r1w_train_launch_summary.txt", "w").write("\n".join(out) + "\n")
print("\n".join(out))
PY
