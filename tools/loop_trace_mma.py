"""MMA-issuer timeline from a DCAP_LOOP_TRACE dump: per pair and step, time spent issuing MMAs, waiting for a free
accumulator buffer (epilogue late) and waiting for operands (producer / dependency late).
Usage: python tools/loop_trace_mma.py trace.bin"""
import sys
import numpy as np
raw = open(sys.argv[1], 'rb').read()
hdr = np.frombuffer(raw[:64], dtype=np.int32)
pairs, items, P, ips = (int(x) for x in hdr[:4])
first = [0] + [int(x) for x in hdr[4:8]]
tiles_m = int(hdr[8]); skew = np.array([0] + [int(x) for x in hdr[9:13]])
d = np.frombuffer(raw[64:], dtype=np.uint64).reshape(pairs, items, 12).astype(np.float64)
t0 = d[d > 0].min(); d = np.where(d > 0, (d - t0) / 1e3, np.nan)
idx = np.arange(items)[None, :] * pairs + np.arange(pairs)[:, None]
j = idx % ips
stage = (j >= first[1]).astype(int) + (j >= first[2]) + (j >= first[3]) + (j >= first[4])
u = idx // ips - skew[stage]
live = (u >= 0) & (u < P * tiles_m)
step = u // tiles_m
steps = list(range(4, min(P, 10)))
tot = dict(mma=0.0, tmemwait=0.0, opwait=0.0); by = {s: dict(tmemwait=0.0, opwait=0.0, n=0) for s in range(4)}
for pr in range(pairs):
    seq = [n for n in range(items) if live[pr, n] and stage[pr, n] < 4 and step[pr, n] in steps]
    for a, b in zip(seq[:-1], seq[1:]):
        e5, m3, m4, m5 = d[pr, a, 5], d[pr, b, 3], d[pr, b, 4], d[pr, b, 5]
        if np.isnan([e5, m3, m4, m5]).any(): continue
        tot['tmemwait'] += m3 - e5; tot['opwait'] += m4 - m3; tot['mma'] += m5 - m4
        s = stage[pr, b]; by[s]['tmemwait'] += m3 - e5; by[s]['opwait'] += m4 - m3; by[s]['n'] += 1
k = pairs * len(steps)
print({a: round(v / k, 1) for a, v in tot.items()}, 'us per pair and step')
for s in range(4):
    n = max(by[s]['n'], 1)
    print('stage %d: per item tmem wait %.2f us, operand wait %.2f us' % (s, by[s]['tmemwait'] / n, by[s]['opwait'] / n))
