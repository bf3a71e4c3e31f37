"""Per-wave kernel durations (device-side events inside the wave graph) for the bench workload."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from __graft_entry__ import load_package
load_package()
from comfyui_distributed_b200 import engine, planner
from comfyui_distributed_b200.denoise import T0Denoiser
B, H, W = 1, 4320, 7680
img = torch.rand(B, H, W, 3, device="cuda")
den = T0Denoiser(123, 0.5)
prof = engine.KernelProfile()
engine.PROFILE = prof
for _ in range(3):
    prof.begin_step()
    engine.upscale_single(img, den, 512, 512, 32, 8, True)
torch.cuda.synchronize()
plan = planner.get_plan(W, H, 512, 512, 32, 8, True)
waves = plan.waves()
recs = prof.graph_rec
crop = [(e0.elapsed_time(e1) * 1e3, nb) for n, e0, e1, nb in recs if n == "crop_resize"]
blend = [(e0.elapsed_time(e1) * 1e3, nb) for n, e0, e1, nb in recs if n == "blend"]
gaps = []
for i in range(len(recs) - 1):
    gaps.append(recs[i][2].elapsed_time(recs[i + 1][1]) * 1e3)
print("wave tiles crop_us crop_GBs blend_us blend_GBs  gap_after_crop(denoise)_us")
for i, w in enumerate(waves):
    print(f"{i:3d} {len(w):3d} {crop[i][0]:8.1f} {crop[i][1]/crop[i][0]/1e3:8.0f} {blend[i][0]:8.1f} {blend[i][1]/blend[i][0]/1e3:8.0f} {gaps[2*i]:8.1f}")
# the same job without per-kernel event nodes in the graph
engine.PROFILE = None
for _ in range(3):
    engine.upscale_single(img, den, 512, 512, 32, 8, True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record()
for _ in range(10):
    engine.upscale_single(img, den, 512, 512, 32, 8, True)
e1.record(); torch.cuda.synchronize()
print("ms/step without event nodes:", e0.elapsed_time(e1) / 10)
engine.PROFILE = prof
for _ in range(2):
    prof.begin_step(); engine.upscale_single(img, den, 512, 512, 32, 8, True)
torch.cuda.synchronize(); e0.record()
for _ in range(10):
    prof.begin_step(); engine.upscale_single(img, den, 512, 512, 32, 8, True)
e1.record(); torch.cuda.synchronize()
print("ms/step with event nodes:", e0.elapsed_time(e1) / 10)
print("sum crop", sum(c[0] for c in crop), "sum blend", sum(b[0] for b in blend), "sum gaps", sum(gaps))
