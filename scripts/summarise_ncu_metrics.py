"""ncu --csv (--metrics ...) log -> one line per (kernel, launch index): duration, dram bytes, L2 hit rate.  Usage: summarise_ncu_metrics.py log.csv"""
import csv, sys, collections, io

rows = []
with open(sys.argv[1], newline="") as f:
    text = f.read()
start = text.find('"ID"')
rd = csv.DictReader(io.StringIO(text[start:]))
per = collections.OrderedDict()
for r in rd:
    key = (int(r["ID"]), r["Kernel Name"])
    per.setdefault(key, {})[r["Metric Name"]] = (r["Metric Value"].replace(",", ""), r["Metric Unit"])
print(f"{'id':>4} {'kernel':60s} {'time_us':>10} {'dram_rd_MB':>11} {'dram_wr_MB':>11} {'GB/s':>8} {'l2_hit%':>8} {'dram%':>7}")
for (i, k), m in per.items():
    def val(name, scale=1.0):
        v = m.get(name)
        if not v:
            return float("nan")
        x, u = float(v[0]), v[1]
        mult = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9, "%": 1.0, "": 1.0}.get(u, 1.0)
        return x * mult * scale
    t = val("gpu__time_duration.sum")            # ns
    rdb, wrb = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
    gbs = (rdb + wrb) / t if t == t and t > 0 else float("nan")
    print(f"{i:4d} {k[:60]:60s} {t / 1e3:10.1f} {rdb / 1e6:11.2f} {wrb / 1e6:11.2f} {gbs:8.1f} {val('lts__t_sector_hit_rate.pct'):8.1f} {val('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):7.1f}")
