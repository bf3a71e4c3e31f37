"""Where does the host->host time go?  PCIe copies alone / concurrently, and the band pipeline."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from __graft_entry__ import load_package
load_package()
from comfyui_distributed_b200 import engine
from comfyui_distributed_b200.denoise import T0Denoiser
B, H, W = 1, 4320, 7680
host = torch.rand(B, H, W, 3).pin_memory()
out = torch.empty_like(host).pin_memory()
dev = torch.empty(B, H, W, 3, device="cuda")
dev2 = torch.rand(B, H, W, 3, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, n=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3
def h2d(): dev.copy_(host, non_blocking=True)
def d2h(): out.copy_(dev2, non_blocking=True)
def both():
    with torch.cuda.stream(s1): dev.copy_(host, non_blocking=True)
    with torch.cuda.stream(s2): out.copy_(dev2, non_blocking=True)
def banded():
    for k in range(8):
        a, b = k * 540, (k + 1) * 540
        with torch.cuda.stream(s1): dev[:, a:b].copy_(host[:, a:b], non_blocking=True)
        with torch.cuda.stream(s2): out[:, a:b].copy_(dev2[:, a:b], non_blocking=True)
print("H2D alone ms", t(h2d), " D2H alone ms", t(d2h), " both concurrently ms", t(both), " both in 8 bands ms", t(banded))
den = T0Denoiser(123, 0.5)
for nb in (1, 2, 4, 9):
    f = lambda: engine.upscale_host(host, den, 512, 512, 32, 8, True, n_bands=nb)
    print("bands", nb, "ms", t(f, 5))
t0 = time.perf_counter(); x = torch.empty(host.shape, dtype=torch.float32, pin_memory=True); print("pinned alloc ms", (time.perf_counter() - t0) * 1e3)
del x
t0 = time.perf_counter(); x = torch.empty(host.shape, dtype=torch.float32, pin_memory=True); print("pinned alloc again ms", (time.perf_counter() - t0) * 1e3)
from comfyui_distributed_b200.nodes import UltimateSDUpscaleDistributed
from comfyui_distributed_b200.testing import T0Model
node = UltimateSDUpscaleDistributed(); model = T0Model()
g = lambda: node.run(host, model, None, None, None, 123, 20, 8.0, "euler", "normal", 0.5, 512, 512, 32, 8, True, False)[0]
print("node.run (result dropped) ms", t(g, 5))
keep = None
def h():
    global keep
    keep = g()
print("node.run (result kept, rebinding) ms", t(h, 8))
lst=[]
def h2():
    lst.append(g())
    if len(lst) > 2: lst.pop(0)
print("node.run (two results kept) ms", t(h2, 8))
import comfyui_distributed_b200.engine as E
print("pool sizes", {k: len(v) for k, v in E.PINNED_RESULTS.bufs.items()})
import cProfile, pstats
pr = cProfile.Profile(); pr.enable(); h(); torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(12)
