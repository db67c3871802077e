"""Top stall-sample SASS lines per kernel from `ncu -i X.ncu-rep --page source --csv` output.
usage: ncu -i rep --page source --csv | python tools/ncu_hot.py [topN] [kernel-substring]"""
import csv, sys
top = int(sys.argv[1]) if len(sys.argv) > 1 else 25
want = sys.argv[2] if len(sys.argv) > 2 else ""
rows = list(csv.reader(sys.stdin))
kern, hdr, body, seen = None, None, [], 0
def flush():
    global body
    if kern is None or not body or (want and want not in kern):
        body = []; return
    si = hdr.index("# Samples"); ii = hdr.index("Instructions Executed")
    tot = sum(int(r[si] or 0) for r in body) or 1
    print("=== %s  (samples %d)" % (kern[:110], tot))
    idx = sorted(range(len(body)), key=lambda i: -int(body[i][si] or 0))[:top]
    for i in sorted(idx):
        r = body[i]
        print("%6d %5.1f%% x%-8s %s" % (i, 100.0 * int(r[si] or 0) / tot, r[ii], r[1].strip()[:110]))
    body = []
for r in rows:
    if len(r) >= 2 and r[0] == "Kernel Name":
        flush(); kern = r[1]; hdr = None; continue
    if r and r[0] == "Address":
        hdr = r; continue
    if hdr is not None and len(r) > 5:
        body.append(r)
flush()
