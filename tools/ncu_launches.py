"""Per-kernel totals from an `ncu --metrics ... --csv --log-file X.csv` launch list.
usage: python tools/ncu_launches.py launches.csv"""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
h = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
hdr = rows[h]
ki, mn, mi, idi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
per = collections.OrderedDict()
for r in rows[h + 1:]:
    if len(r) <= mi:
        continue
    per.setdefault((int(r[idi]), r[ki][:70]), {})[r[mn]] = float(r[mi].replace(",", ""))
agg, cnt = collections.defaultdict(lambda: collections.defaultdict(float)), collections.Counter()
for (i, n), m in per.items():
    cnt[n] += 1
    for k, v in m.items():
        agg[n][k] += v
T = "gpu__time_duration.sum"
tot = sum(a[T] for a in agg.values())
print("launches %d, total %.1f us (cold-cache, serialised: compare shares)" % (len(per), tot / 1e3))
for n, a in sorted(agg.items(), key=lambda x: -x[1][T]):
    c = cnt[n]
    extra = ""
    if "dram__bytes_read.sum" in a:
        extra = " rd %7.1f MB wr %7.1f MB" % (a["dram__bytes_read.sum"] / c / 1e6, a["dram__bytes_write.sum"] / c / 1e6)
    tp = "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"
    if tp in a:
        extra += " tensor %5.1f%%" % (a[tp] / c)
    print("%9.1f us %5.1f%% x%4d avg %8.1f us%s  %s" % (a[T] / 1e3, 100 * a[T] / tot, c, a[T] / c / 1e3, extra, n))
