"""Where a kernel's instructions and stall samples go, by source line (needs -lineinfo and --import-source on):
python tools/ncu_regions.py report.ncu-rep [top]"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr, agg, cur = None, collections.OrderedDict(), None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        iI, iS = hdr.index("Instructions Executed"), hdr.index("# Samples")
        continue
    if hdr is None:
        continue
    if r[0].isdigit():
        key = (cur, int(r[0]), r[1].strip()[:100])
        try:
            a = agg.setdefault(key, [0, 0])
            a[0] += int(r[iI]); a[1] += int(r[iS])
        except Exception:
            pass
tot = sum(v[0] for v in agg.values()); tots = sum(v[1] for v in agg.values())
print(f"warp instructions {tot}  samples {tots}")
print("--- by samples")
for (f, ln, src), (n, s) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"samples {s / tots:6.3f} instr {n / tot:6.3f}  {f}:{ln:<4d} {src}")
