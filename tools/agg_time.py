"""Aggregate DAD_DEBUG_TIME=1 per-launch lines (stderr) by label: python tools/agg_time.py file [marker]"""
import re, sys, collections
txt = open(sys.argv[1]).read()
if len(sys.argv) > 2:
    txt = txt.split(sys.argv[2])[-1]
agg = collections.defaultdict(lambda: [0, 0.0])
for l in txt.splitlines():
    m = re.match(r"dad\[time\]\s+([\d.]+) ms\s+(.*)", l)
    if m:
        lab = re.sub(r"\d+", "#", m.group(2))[:60]
        agg[lab][0] += 1
        agg[lab][1] += float(m.group(1))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print("%9.2f ms %5d  %s" % (v[1], v[0], k))
print("total", sum(v[1] for v in agg.values()))
