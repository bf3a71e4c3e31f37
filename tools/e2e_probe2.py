import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from __graft_entry__ import load_package
load_package()
from comfyui_distributed_b200 import engine as E
from comfyui_distributed_b200.denoise import T0Denoiser
B, H, W = 1, 4320, 7680
host = torch.rand(B, H, W, 3).pin_memory()
den = T0Denoiser(123, 0.5)
plan = E.get_plan(W, H, 512, 512, 32, 8, True)
dp = E.DevicePlan.get(plan, torch.device("cuda", 0))
hp = E.HostPipeline.get(dp, B, den, 4)
bufs = [torch.empty(host.shape, pin_memory=True) for _ in range(3)]
for b in bufs: hp.run(host, b); torch.cuda.synchronize()
def run(seq, label):
    ts = []
    for i in seq:
        torch.cuda.synchronize(); t0 = time.perf_counter()
        hp.run(host, bufs[i]); t1 = time.perf_counter()
        torch.cuda.synchronize(); t2 = time.perf_counter()
        ts.append(((t1 - t0) * 1e3, (t2 - t0) * 1e3))
    print(label, " ".join(f"{a:.1f}/{b:.1f}" for a, b in ts))
run([0, 0, 0, 0, 0, 0], "same buffer      enqueue/total ms:")
run([0, 1, 0, 1, 0, 1], "alternating 0/1  enqueue/total ms:")
run([0, 1, 2, 0, 1, 2], "rotating 0/1/2   enqueue/total ms:")
# touch the buffer on the CPU between runs (what a consumer would do)
for i in [0, 1, 0, 1]:
    torch.cuda.synchronize(); t0 = time.perf_counter(); hp.run(host, bufs[i]); torch.cuda.synchronize(); t1 = time.perf_counter()
    s = float(bufs[i][0, ::64, ::64].sum()); print("alt + cpu read: %.1f ms" % ((t1 - t0) * 1e3))
