"""`ncu -i X.ncu-rep --page raw --csv` -> one line per captured launch: duration, DRAM bytes read + written, achieved DRAM GB/s and its share of
the measured HBM peak (MEASURED_PEAKS.json), tensor-pipe active %, issue-active %, registers, occupancy.  Also writes the per-launch DRAM traffic
of the dominant kernel family (convgemm) to profiles/r02_ncu_traffic.json when --traffic-json is given.
Usage: python tools/ncu_summarize.py raw.csv "<profiled command>" [--traffic-json path workload_key_json]"""
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rows = list(csv.reader(open(sys.argv[1], errors="replace")))
hdr = next(r for r in rows if "Kernel Name" in r)
i0 = rows.index(hdr)
units = rows[i0 + 1]
col = {n: i for i, n in enumerate(hdr)}
try:
    hbm = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    src = "measured (MEASURED_PEAKS.json)"
except Exception:
    hbm, src = 6650.0, "fallback (B200_PROFILING.md)"


def val(r, name, default=0.0):
    i = col.get(name)
    if i is None or i >= len(r) or r[i] in ("", "n/a"):
        return default
    return float(r[i].replace(",", ""))


def to_us(r):
    v, u = val(r, "gpu__time_duration.sum"), units[col["gpu__time_duration.sum"]]
    return v * {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3}.get(u, 1e-3)


def to_bytes(r, name):
    v, u = val(r, name), units[col[name]] if name in col else "byte"
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)


print(f"# ncu --set full --clock-control none (B200, sm_100a), one launch per kernel instance -- command: {sys.argv[2] if len(sys.argv) > 2 else ''}")
print(f"# HBM peak for the %% column: {hbm} GB/s, {src}.  ncu serialises and replays launches (cold caches): durations are NOT bench numbers.")
print(f"{'kernel':58s} {'grid':>7s} {'us':>8s} {'DRAM rd MB':>10s} {'wr MB':>8s} {'GB/s':>7s} {'%HBM':>6s} {'tensor%':>8s} {'issue%':>7s} {'regs':>5s} {'warps%':>7s}")
conv_traffic = []
for r in rows[i0 + 2:]:
    if len(r) <= col["Kernel Name"]:
        continue
    name = re.sub(r"\(.*", "", r[col["Kernel Name"]]).replace("(anonymous namespace)::", "").replace("<unnamed>::", "").replace("void ", "")
    us = to_us(r)
    rd, wr = to_bytes(r, "dram__bytes_read.sum"), to_bytes(r, "dram__bytes_write.sum")
    gbs = (rd + wr) / (us * 1e-6) / 1e9 if us > 0 else 0
    grid = r[col["Grid Size"]] if "Grid Size" in col else ""
    print(f"{name[:58]:58s} {grid:>7s} {us:8.1f} {rd / 1e6:10.2f} {wr / 1e6:8.2f} {gbs:7.0f} {100 * gbs / hbm:6.1f} "
          f"{val(r, 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', val(r, 'sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active')):8.1f} "
          f"{val(r, 'sm__issue_active.avg.pct_of_peak_sustained_active', val(r, 'smsp__issue_active.avg.pct_of_peak_sustained_active')):7.1f} "
          f"{int(val(r, 'launch__registers_per_thread')):5d} {val(r, 'sm__warps_active.avg.pct_of_peak_sustained_active'):7.1f}")
    if name.startswith("convgemm_kernel"):
        conv_traffic.append(rd + wr)
if "--traffic-json" in sys.argv and conv_traffic:
    k = sys.argv.index("--traffic-json")
    out = {"workload_key": json.loads(sys.argv[k + 2]), "avg_dram_bytes_per_launch": sum(conv_traffic) / len(conv_traffic), "launches": len(conv_traffic),
           "note": f"mean of dram__bytes_read.sum + dram__bytes_write.sum over the {len(conv_traffic)} convgemm_kernel launches of tools/ncu_kernels.py (the UNet / VAE shapes of the "
                   "bench workload), ncu --set full --clock-control none; per-launch values in profiles/r02_ncu_kernels.txt"}
    json.dump(out, open(sys.argv[k + 1], "w"), indent=1)
