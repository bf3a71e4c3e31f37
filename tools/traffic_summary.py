"""Summarise an `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --csv` launch list of
tools/one_step.py into profiles/<name>.json: DRAM traffic of the seam-blend (and crop) launches of ONE step, next to the
algorithmic bytes, stamped with the hash of the kernel sources it was captured from (bench.py prints `roofline.traffic`
only when that hash matches the sources it runs).
usage: python tools/traffic_summary.py launches.csv out.json [workload] [source-hash file written on the GPU box beside the
capture: `python -c "import bench; print(bench.kernel_source_hash())" > hash.txt`]"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench

src, out = sys.argv[1], sys.argv[2]
workload = sys.argv[3] if len(sys.argv) > 3 else "cfg2_4k_to_8k_sdxl_512px"
rows = list(csv.DictReader([l for l in open(src) if l.startswith('"')]))
launches = {}
for r in rows:
    launches.setdefault(int(r["ID"]), {"name": r["Kernel Name"], "grid": r["Grid Size"]})[r["Metric Name"]] = (float(r["Metric Value"]), r["Metric Unit"])


def to_bytes(v):
    val, unit = v
    return val * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def to_us(v):
    val, unit = v
    return val * {"ns": 1e-3, "us": 1, "ms": 1e3, "usecond": 1, "nsecond": 1e-3, "msecond": 1e3}.get(unit, 1)


from __graft_entry__ import load_package  # noqa: E402

load_package()
from comfyui_distributed_b200 import planner  # noqa: E402

B, H, W, tile, pad, blur = bench.WORKLOADS[workload]
plan = planner.get_plan(W, H, tile, tile, pad, blur, True)
n_waves = len(plan.waves())
src_hash = open(sys.argv[4]).read().strip() if len(sys.argv) > 4 else bench.kernel_source_hash()
res = {"workload": workload, "source_hash": src_hash, "launches_per_step": n_waves,
       "how": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none on tools/one_step.py; "
              "the LAST launches_per_step launches of each kernel = the timed step (cold-cache, serialised: bytes are meaningful, times are not)"}
for key, pat in (("blend", "blend_"), ("crop_resize", "crop_")):
    ls = [v for k, v in sorted(launches.items()) if pat in v["name"]][-n_waves:]
    rd = sum(to_bytes(v["dram__bytes_read.sum"]) for v in ls)
    wr = sum(to_bytes(v["dram__bytes_write.sum"]) for v in ls)
    res[key] = {"kernel": ls[0]["name"].split("(")[0] if ls else None, "launches": len(ls), "dram_read_bytes_per_step": rd,
                "dram_write_bytes_per_step": wr, "dram_bytes_per_step": rd + wr, "dram_bytes_per_launch": (rd + wr) / max(len(ls), 1),
                "sum_duration_us_under_ncu": round(sum(to_us(v["gpu__time_duration.sum"]) for v in ls), 1)}
alg = {"crop_resize": 0, "blend": 0}
for w in plan.waves():
    cw, offs_w, _ = plan.crop_worklist(w, B)
    alg["crop_resize"] += cw.algo_bytes * B
    alg["blend"] += plan.blend_worklist(w, offs_w, 4, None, B).algo_bytes * B
for k in alg:
    res[k]["algorithmic_bytes_per_step"] = alg[k]
    res[k]["traffic_over_algorithmic"] = round(res[k]["dram_bytes_per_step"] / alg[k], 3)
json.dump(res, open(out, "w"), indent=1)
print(json.dumps(res, indent=1))
