"""Per-kernel counts of the SASS mnemonics that prove a Blackwell-native kernel (B200_PROFILING.md: tcgen05.mma -> UTC*MMA, tcgen05.ld/st ->
LDTM/STTM, TMA -> UTMALDG/UTMASTG; legacy mma.sync would show as HMMA) in the in-tree libladi_b200.so.  No GPU needed:
    python tools/sass_grep.py > profiles/r02_sass_grep.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "ladi_vton_b200", "libladi_b200.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
pats = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "HMMA", "MUFU.EX2", "MUFU.TANH", "SYNCS", "BAR.SYNC", "BAR.ARV"]
rows, cur = collections.OrderedDict(), None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(anonymous namespace\)::", "", name)
        name = re.sub(r"\(.*", "", name).replace("void ", "")
        cur = rows.setdefault(name, collections.Counter())
        continue
    if cur is None:
        continue
    for p in pats:
        if re.search(r"\b" + re.escape(p) + r"\b", line) or (p.startswith("MUFU") and p in line):
            cur[p] += 1
    if re.match(r"\s+/\*[0-9a-f]{4}\*/", line):
        cur["instructions"] += 1
print(f"# cuobjdump -sass {os.path.relpath(so, ROOT)} | per-kernel mnemonic counts (tools/sass_grep.py); sm_100a, CUDA {subprocess.run(['nvcc', '--version'], capture_output=True, text=True).stdout.split('release ')[-1].split(',')[0]}")
print(f"{'kernel':70s} " + " ".join(f"{p:>9s}" for p in pats) + f" {'instr':>8s}")
for name, c in rows.items():
    print(f"{name[:70]:70s} " + " ".join(f"{c[p]:9d}" for p in pats) + f" {c['instructions']:8d}")
tot = collections.Counter()
for c in rows.values():
    tot.update(c)
print(f"{'TOTAL':70s} " + " ".join(f"{tot[p]:9d}" for p in pats) + f" {tot['instructions']:8d}")
assert tot["HMMA"] == 0, "legacy mma.sync (HMMA) found"
