import sys, os
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/oracle'); sys.path.insert(0, '/root/repo/tests')
import numpy as np, torch
import usdu_oracle as orc
from __graft_entry__ import load_package
from inputs import make_input
load_package()
from comfyui_distributed_b200 import engine, planner, _native as nat
kind,B,H,W,tile,pad,uniform = "noise",1,333,777,64,128,True
img = make_input(kind, 3, B, H, W)
p = planner.Plan.build(W, H, tile, tile, pad, 8, uniform)
dp = engine.DevicePlan.get(p, torch.device("cuda:0"))
canvas = engine.Canvas(dp, B).load(torch.from_numpy(img).cuda())
ids = list(range(len(p.tiles)))
buf, offs = canvas.crop(ids)
cu8 = orc.quantize_u8(img)
oplan = orc.make_plan(W, H, tile, tile, pad, uniform)[2]
host = buf.cpu().numpy()
bad = 0
for i, t in enumerate(oplan):
    ref = orc.extract_tile(cu8, t)
    got = host[offs[i]: offs[i] + ref.size].reshape(ref.shape)
    if not np.array_equal(got, ref):
        d = (got != ref)[0].any(-1)
        ys, xs = np.nonzero(d)
        pt = p.tiles[i]
        print("tile", i, "ew,eh", t.ew, t.eh, "pw,ph", t.pw, t.ph, "taps", p._tab_taps[(t.ew,t.pw)], p._tab_taps[(t.eh,t.ph)], "bad px", d.sum(), "rows", ys.min(), ys.max(), "cols", xs.min(), xs.max())
        bad += 1
        if bad > 6: break
print("bad tiles", bad, "of", len(oplan))
