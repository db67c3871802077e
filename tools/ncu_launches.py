"""Per-kernel totals from an `ncu --metrics ... --csv --log-file X.csv` launch list.
usage: python tools/ncu_launches.py launches.csv [--last N] [--list]
  --last N   only the last N launches (e.g. one step)
  --list     also print every launch in order (id, time, grid, kernel)"""
import csv, collections, sys
args = sys.argv[1:]
path = args[0]
last = int(args[args.index("--last") + 1]) if "--last" in args else 0
rows = list(csv.reader(open(path)))
h = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
hdr = rows[h]
ki, mn, mi, idi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
gi = hdr.index("Grid Size") if "Grid Size" in hdr else None
per = collections.OrderedDict()
for r in rows[h + 1:]:
    if len(r) <= mi:
        continue
    key = (int(r[idi]), r[ki][:90], r[gi] if gi is not None else "")
    per.setdefault(key, {})[r[mn]] = float(r[mi].replace(",", ""))
items = list(per.items())
if last:
    items = items[-last:]
T = "gpu__time_duration.sum"
tp = "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"
if "--list" in args:
    for (i, n, g), m in items:
        ex = ""
        if "dram__bytes_read.sum" in m:
            ex = " rd %7.1f wr %7.1f MB" % (m["dram__bytes_read.sum"] / 1e6, m["dram__bytes_write.sum"] / 1e6)
        if tp in m:
            ex += " tensor %5.1f%%" % m[tp]
        print("%6d %9.1f us%s grid %-14s %s" % (i, m[T] / 1e3, ex, g, n))
agg, cnt = collections.defaultdict(lambda: collections.defaultdict(float)), collections.Counter()
for (i, n, g), m in items:
    cnt[n] += 1
    for k, v in m.items():
        agg[n][k] += v
tot = sum(a[T] for a in agg.values())
print("launches %d, total %.1f us (cold-cache, serialised: compare shares)" % (len(items), tot / 1e3))
for n, a in sorted(agg.items(), key=lambda x: -x[1][T]):
    c = cnt[n]
    extra = ""
    if "dram__bytes_read.sum" in a:
        extra = " rd %7.1f MB wr %7.1f MB" % (a["dram__bytes_read.sum"] / c / 1e6, a["dram__bytes_write.sum"] / c / 1e6)
    if tp in a:
        extra += " tensor %5.1f%%" % (a[tp] / c)
    print("%9.1f us %5.1f%% x%4d avg %8.1f us%s  %s" % (a[T] / 1e3, 100 * a[T] / tot, c, a[T] / c / 1e3, extra, n))
