"""Prints ms_per_step and the roofline fraction(s) of the last JSON line on stdin (helper of the GPU session scripts)."""
import json
import sys

d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print(d.get("ms_per_step"), {k: (d.get(k) or {}).get("frac") for k in ("roofline", "roofline_hbm") if d.get(k)}, d.get("value"))
