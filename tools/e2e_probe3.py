import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from __graft_entry__ import load_package
load_package()
from comfyui_distributed_b200 import engine as E
from comfyui_distributed_b200.nodes import UltimateSDUpscaleDistributed
from comfyui_distributed_b200.testing import T0Model
B, H, W = 1, 4320, 7680
host = torch.rand(B, H, W, 3).pin_memory()
node = UltimateSDUpscaleDistributed(); model = T0Model()
g = lambda: node.run(host, model, None, None, None, 123, 20, 8.0, "euler", "normal", 0.5, 512, 512, 32, 8, True, False)[0]
g(); g(); torch.cuda.synchronize()
keep = None
for i in range(8):
    t0 = time.perf_counter(); r = g(); t1 = time.perf_counter()
    keep = r; t2 = time.perf_counter()
    print(i, "call %.1f ms  rebind %.2f ms  ptr %x" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3, r.data_ptr()), flush=True)
    del r
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for i in range(4):
    keep = g()
pr.disable(); pstats.Stats(pr).sort_stats("tottime").print_stats(8)
