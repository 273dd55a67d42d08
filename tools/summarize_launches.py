"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list (one row per kernel launch) by kernel name:
launches, total ms, share.  Usage: python tools/summarize_launches.py launches.csv "<command line that was profiled>" > summary.txt"""
import csv, re, sys
from collections import defaultdict
rows = list(csv.reader(l for l in open(sys.argv[1], errors="replace") if l.startswith('"')))
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg, cnt = defaultdict(float), defaultdict(int)
for r in rows[1:]:
    if len(r) <= vi:
        continue
    name = re.sub(r"\(.*", "", r[ki]).replace("(anonymous namespace)::", "").replace("<unnamed>::", "")
    name = re.sub(r"void ", "", name)
    v = float(r[vi].replace(",", ""))
    v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0}.get(r[ui], 1e-6)
    agg[name] += v
    cnt[name] += 1
tot = sum(agg.values())
print(f"# ncu launch list (B200, sm_100a) -- command: {sys.argv[2] if len(sys.argv) > 2 else ''}")
print("# per-launch times are cold-cache / serialised (ncu replays each launch): compare SHARES with the in-graph timeline, not absolutes")
print(f"# total launches {sum(cnt.values())}, total kernel time {tot:.1f} ms")
print("launches         ms   share  kernel")
for k, v in sorted(agg.items(), key=lambda x: -x[1]):
    print(f"{cnt[k]:8d} {v:10.2f} {100 * v / tot:6.2f}%  {k}")
