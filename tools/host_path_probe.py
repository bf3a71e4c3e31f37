"""Host-path timing: pinned vs pageable input through node.run (cfg2), per call wall clock."""
import sys, time
import torch
sys.path.insert(0, ".")
from __graft_entry__ import load_package
load_package()
from comfyui_distributed_b200 import engine
from comfyui_distributed_b200.nodes import UltimateSDUpscaleDistributed
from comfyui_distributed_b200.testing import T0Model

torch.cuda.set_device(0)
g = torch.Generator().manual_seed(0)
img = torch.floor(torch.rand(1, 4320, 7680, 3, generator=g) * 255) / 255
node = UltimateSDUpscaleDistributed()
for name, x in (("pinned", img.pin_memory()), ("pageable", img)):
    for bands in (None, 4):
        ts, keep = [], None
        for i in range(8):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            if bands is None:
                (keep,) = node.run(x, T0Model(), None, None, None, 123, 20, 8.0, "euler", "normal", 0.5, 512, 512, 32, 8, True, False)
            else:
                keep = engine.upscale_host(x, T0Model().as_usdu_denoiser(seed=123, denoise=0.5), 512, 512, 32, 8, True, n_bands=bands)
            torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
        print(name, "bands", bands or "auto", " ".join(f"{t:.1f}" for t in ts), flush=True)
# plain copies for scale
d = torch.empty_like(img, device="cuda")
for name, x in (("pinned", img.pin_memory()), ("pageable", img)):
    for _ in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter(); d.copy_(x, non_blocking=True); torch.cuda.synchronize()
        print("H2D", name, f"{(time.perf_counter() - t0) * 1e3:.1f} ms")
st = torch.empty_like(img).pin_memory()
for _ in range(3):
    t0 = time.perf_counter(); st.copy_(img); print("host memcpy to pinned", f"{(time.perf_counter() - t0) * 1e3:.1f} ms", "threads", torch.get_num_threads())
