"""One device-resident step of the bench workload (after warm-up), for captures: `ncu ... python tools/one_step.py [workload]`."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from __graft_entry__ import load_package

load_package()
import bench
from comfyui_distributed_b200 import engine
from comfyui_distributed_b200.denoise import T0Denoiser

name = sys.argv[1] if len(sys.argv) > 1 else "cfg2_4k_to_8k_sdxl_512px"
B, H, W, tile, pad, blur = bench.WORKLOADS[name]
img = bench.make_canvas_cpu(B, H, W).cuda()
den = T0Denoiser(bench.SEED, bench.DENOISE)
for _ in range(2):
    engine.upscale_single(img, den, tile, tile, pad, blur, True)
torch.cuda.synchronize()
print("STEP-BEGIN", flush=True)
engine.upscale_single(img, den, tile, tile, pad, blur, True)
torch.cuda.synchronize()
