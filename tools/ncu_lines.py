#!/usr/bin/env python
"""Per-source-line executed instructions / stall samples of one kernel in an .ncu-rep.
usage: python tools/ncu_lines.py rep kernel_regex [top]"""
import csv, subprocess, sys, collections
rep, kre = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", f"regex:{kre}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
cur_file = None
agg = collections.OrderedDict()
seen_kernels = 0
hdr = None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = r
        iI, iS = hdr.index("Instructions Executed"), hdr.index("# Samples")
        continue
    if hdr is None:
        continue
    if r[0] != "" and r[0].isdigit():       # a source line summary row
        key = (cur_file, int(r[0]), r[1].strip()[:90])
        try:
            a = agg.setdefault(key, [0, 0])
            a[0] += int(r[iI]); a[1] += int(r[iS])
        except Exception:
            pass
tot = sum(v[0] for v in agg.values()); tots = sum(v[1] for v in agg.values())
print(f"total warp instr {tot}  samples {tots}")
for (f, ln, src), (n, s) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{n/tot:6.3f} {s/max(tots,1):6.3f}  {f}:{ln:<4d} {src}")
