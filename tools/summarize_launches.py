"""Sum an `ncu --csv --metrics ...` launch list per kernel: n, total ms, share, DRAM MB, DRAM GB/s.
    python tools/summarize_launches.py gpurun_out/launches.csv [first_launch_id]"""
import csv, re, sys, collections
rows = collections.defaultdict(dict)
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    rows[int(r["ID"])]["name"] = re.sub(r"\(.*", "", r["Kernel Name"].replace("<unnamed>::", "").replace("void ", ""))
    try:
        rows[int(r["ID"])][r["Metric Name"]] = (float(r["Metric Value"].replace(",", "")), r["Metric Unit"])
    except ValueError:
        pass
first = int(sys.argv[2]) if len(sys.argv) > 2 else 0
agg = collections.OrderedDict()
def to_bytes(v):
    val, unit = v
    return val * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
def to_ns(v):
    val, unit = v
    return val * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9, "nsecond": 1, "usecond": 1e3, "msecond": 1e6, "second": 1e9}.get(unit, 1)
for i in sorted(rows):
    if i < first: continue
    r = rows[i]; a = agg.setdefault(r["name"], [0, 0.0, 0.0, 0.0])
    a[0] += 1; a[1] += to_ns(r.get("gpu__time_duration.sum", (0, "ns")))
    a[2] += to_bytes(r.get("dram__bytes_read.sum", (0, "byte"))); a[3] += to_bytes(r.get("dram__bytes_write.sum", (0, "byte")))
tot = sum(a[1] for a in agg.values())
print("total %.1f ms over %d launches" % (tot * 1e-6, sum(a[0] for a in agg.values())))
print("%-34s %5s %9s %6s %10s %10s %9s" % ("kernel", "n", "ms", "share", "dram_rd_MB", "dram_wr_MB", "dram_GB/s"))
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-34s %5d %9.3f %5.1f%% %10.1f %10.1f %9.1f" % (k[:34], a[0], a[1] * 1e-6, 100 * a[1] / tot, a[2] * 1e-6, a[3] * 1e-6, (a[2] + a[3]) / a[1] if a[1] else 0))
