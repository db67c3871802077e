"""Summary of a DCAP_LOOP_TRACE dump (csrc/greedy_loop.cu): per stage, how long the leader CTA of a pair spends in
each phase of an item.  Marks: 0 producer reaches the item, 1 dependency met, 2 proxy fence done, 3 MMA issuer has the
accumulator buffer, 4 first operands landed, 5 last MMA issued, 6 epilogue warp 2 released TMEM, 7 published, 8 epilogue warp 2 starts the item, 9 its body is done (before the TMEM release).
Usage: python tools/loop_trace.py gpurun_out/trace.bin [step]"""
import sys

import numpy as np


def main():
    path = sys.argv[1]
    raw = open(path, "rb").read()
    hdr = np.frombuffer(raw[:64], dtype=np.int32)
    pairs, items, P, ips = (int(x) for x in hdr[:4])
    first = [0] + [int(x) for x in hdr[4:8]]
    tiles_m = int(hdr[8])
    skew = np.array([0] + [int(x) for x in hdr[9:13]])
    d = np.frombuffer(raw[64:], dtype=np.uint64).reshape(pairs, items, 12).astype(np.float64)
    t0 = d[d > 0].min()
    d = np.where(d > 0, (d - t0) / 1e3, np.nan)         # us
    total = (P * tiles_m + int(skew[4])) * ips
    print("pairs %d, items/pair %d, steps %d, items/slot %d, tiles_m %d, skews %s, kernel span %.1f us" % (pairs, items, P, ips, tiles_m, skew.tolist(), np.nanmax(d)))
    idx = np.arange(items)[None, :] * pairs + np.arange(pairs)[:, None]
    j = idx % ips
    stage = (j >= first[1]).astype(int) + (j >= first[2]) + (j >= first[3]) + (j >= first[4])
    u = idx // ips - skew[stage]
    valid = (idx < total) & (u >= 0) & (u < P * tiles_m) & ~np.isnan(d[..., 7])
    step = u // tiles_m
    sel_step = int(sys.argv[2]) if len(sys.argv) > 2 else None
    names = ["dep wait (1-0)", "proxy fence (2-1)", "acc buffer->first operands (4-3)", "mma issue (5-4)",
             "mma end->tmem released (6-5)", "publish (+merge) (7-6)", "item period (5 - prev 5)",
             "epilogue: start->body done (9-8)", "epilogue: own period (8 - prev 8)", "epilogue: start lag behind mma end (8-5)",
             "epilogue: accumulator ready -> first chunk done / first TMEM load back (11-10)", "epilogue: accumulator ready -> body done (9-10)",
             "epilogue: mma end -> accumulator seen (10-5)"]
    for s in range(5):
        m = valid & (stage == s)
        if sel_step is not None:
            m &= step == sel_step
        rows = [d[..., 1] - d[..., 0], d[..., 2] - d[..., 1], d[..., 4] - d[..., 3], d[..., 5] - d[..., 4],
                d[..., 6] - d[..., 5], d[..., 7] - d[..., 6]]
        period = np.full_like(d[..., 5], np.nan)
        period[:, 1:] = d[:, 1:, 5] - d[:, :-1, 5]
        rows.append(period)
        rows.append(d[..., 9] - d[..., 8])
        ep = np.full_like(d[..., 8], np.nan)
        ep[:, 1:] = d[:, 1:, 8] - d[:, :-1, 8]
        rows.append(ep)
        rows.append(d[..., 8] - d[..., 5])
        rows.append(d[..., 11] - d[..., 10])
        rows.append(d[..., 9] - d[..., 10])
        rows.append(d[..., 10] - d[..., 5])
        print("stage %d: %d items" % (s, int(m.sum())))
        for name, r in zip(names, rows):
            v = r[m]
            v = v[~np.isnan(v)]
            if v.size:
                print("   %-36s mean %7.2f  p50 %7.2f  p90 %7.2f  max %8.2f us" % (name, v.mean(), np.median(v), np.percentile(v, 90), v.max()))
    # per-step wall: first mark-0 of the step to last mark-7
    for t in range(P):
        m = valid & (step == t)
        a = np.nanmin(np.where(m, d[..., 0], np.nan))
        b = np.nanmax(np.where(m, d[..., 7], np.nan))
        busy = np.nansum(np.where(m, d[..., 5] - d[..., 4], 0.0)) / pairs
        print("step %2d: %8.1f .. %8.1f us (%.1f), mma-issue time per pair %.1f us" % (t, a, b, b - a, busy))


if __name__ == "__main__":
    main()
