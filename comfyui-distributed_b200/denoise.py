"""Sampler adapters for the per-tile denoise step (upscale/tile_ops.py:239-287).

The tile pipeline treats the sampler as an injected callable on device tensors
(``engine.Denoiser``).  Three are provided:

* ``T0Denoiser``      deterministic stand-in used by the parity tests and the benchmark:
                      x' = clamp(x*(1-d) + noise(seed)*d).  The same seed for every tile,
                      as the reference passes ``seed`` unchanged (tile_ops.py:430-431).
* ``SyntheticSDXL``   a torch conv/attention stack with SDXL-like cost per 544x544 tile
                      (no weights available offline; BASELINE.md tier T1).
* ``ComfySampler``    the real thing when ComfyUI is importable: VAEEncode ->
                      common_ksampler -> VAEDecode[Tiled], one call per tile position with
                      the whole frame batch, like process_tiles_batch.
"""
from __future__ import annotations

from contextlib import nullcontext
from typing import List

import numpy as np
import torch

from .model_patch import cropped_model_patches


class T0Denoiser:
    cuda_graph_safe = True    # pure elementwise device work on fixed shapes

    def __init__(self, seed: int, denoise: float):
        self.seed = int(seed)
        self.graph_key = ("t0", int(seed), float(np.float32(denoise)))
        # every step individually rounded in fp32 so CPU (oracle) and GPU agree bit-for-bit
        self.d = float(np.float32(denoise))
        self.omd = float(np.float32(1.0) - np.float32(denoise))
        self._noise = {}

    def noise(self, shape, device):
        key = (tuple(shape), str(device))
        if key not in self._noise:
            g = torch.Generator().manual_seed(self.seed)          # CPU generator: identical everywhere
            self._noise[key] = (torch.rand(tuple(shape), generator=g, dtype=torch.float32).to(device) * self.d)
        return self._noise[key]

    def __call__(self, tiles: torch.Tensor, rows: List) -> torch.Tensor:
        nd = self.noise(tiles.shape[1:], tiles.device)            # [B, ph, pw, 3], pre-scaled by d
        if tiles.is_cuda and tiles.is_contiguous() and tiles.dtype == torch.float32 and nd.numel() % 4 == 0:
            from . import _native as nat                          # one fused pass (same rounding as the torch ops)
            out = torch.empty_like(tiles)
            nat.t0_denoise(tiles.data_ptr(), nd.data_ptr(), out.data_ptr(), tiles.numel(), nd.numel(), self.omd,
                           torch.cuda.current_stream().cuda_stream)
            return out
        return torch.clamp(tiles * self.omd + nd, 0.0, 1.0)


class SyntheticSDXL(torch.nn.Module):
    """Random-weight latent denoiser with roughly SDXL-UNet FLOPs per step at 68x68
    latents (ph/8): strided conv encoder, `steps` x 2 (cfg) passes of a conv + attention
    trunk in bf16, conv decoder.  Output is blended with the input so values stay in
    [0,1].  Only for end-to-end cost studies (tier T1); not a parity denoiser."""

    def __init__(self, steps: int = 20, width: int = 640, depth: int = 6, denoise: float = 0.5):
        super().__init__()
        self.steps, self.denoise = steps, denoise
        g = torch.Generator().manual_seed(0)
        self.enc = torch.nn.Conv2d(3, width, 8, stride=8)
        self.blocks = torch.nn.ModuleList()
        for _ in range(depth):
            self.blocks.append(torch.nn.ModuleDict({
                "c1": torch.nn.Conv2d(width, width, 3, padding=1),
                "c2": torch.nn.Conv2d(width, width, 3, padding=1),
                "qkv": torch.nn.Linear(width, 3 * width),
                "proj": torch.nn.Linear(width, width),
            }))
        self.dec = torch.nn.ConvTranspose2d(width, 3, 8, stride=8)
        for p in self.parameters():
            with torch.no_grad():
                p.copy_(torch.randn(p.shape, generator=g) * 0.02)

    @torch.no_grad()
    def forward(self, tiles: torch.Tensor, rows: List) -> torch.Tensor:
        n, B, ph, pw, _ = tiles.shape
        x = tiles.reshape(n * B, ph, pw, 3).permute(0, 3, 1, 2).to(torch.bfloat16)
        h = self.enc(x)
        for _ in range(self.steps * 2):
            for blk in self.blocks:
                h = h + torch.nn.functional.silu(blk["c2"](torch.nn.functional.silu(blk["c1"](h))))
                b, c, hh, ww = h.shape
                t = h.flatten(2).transpose(1, 2)
                q, k, v = blk["qkv"](t).view(b, hh * ww, 3, 10, c // 10).permute(2, 0, 3, 1, 4)
                a = torch.nn.functional.scaled_dot_product_attention(q, k, v)
                h = h + blk["proj"](a.transpose(1, 2).reshape(b, hh * ww, c)).transpose(1, 2).view(b, c, hh, ww)
            h = h * 0.5
        y = torch.sigmoid(self.dec(h).float()).permute(0, 2, 3, 1).reshape(n, B, ph, pw, 3)
        return torch.clamp(tiles * (1 - self.denoise) + y * self.denoise, 0.0, 1.0)


class ComfySampler:
    """process_tiles_batch (upscale/tile_ops.py:239-287) against a live ComfyUI: one
    VAEEncode -> common_ksampler -> VAEDecode per tile position with all B frames."""

    def __init__(self, model, positive, negative, vae, seed, steps, cfg, sampler_name, scheduler, denoise,
                 tiled_decode=False, image_size=None, cond_cropper=None):
        import nodes as comfy_nodes  # ComfyUI's module; raises ImportError outside ComfyUI
        self.n = comfy_nodes
        self.args = (model, positive, negative, vae, seed, steps, cfg, sampler_name, scheduler, denoise)
        # both tiled classes must import, like upscale/tile_ops.py:249-254 (the encode stays non-tiled, :276)
        self.tiled_decode = tiled_decode and hasattr(comfy_nodes, "VAEDecodeTiled") and hasattr(comfy_nodes, "VAEEncodeTiled")
        self.image_size = image_size
        self.cond_cropper = cond_cropper
        try:        # user cancel is polled once per tile like upscale/modes/static.py:326,407,476
            import comfy.model_management as mm
            self._poll_interrupt = mm.throw_exception_if_processing_interrupted
        except ImportError:
            self._poll_interrupt = None

    def __call__(self, tiles: torch.Tensor, rows: List) -> torch.Tensor:
        model, positive, negative, vae, seed, steps, cfg, sampler_name, scheduler, denoise = self.args
        outs = []
        for i, row in enumerate(rows):
            px = tiles[i]
            pos, neg = positive, negative
            if self.cond_cropper is not None:
                pos, neg = self.cond_cropper(positive, negative, row, (px.shape[2], px.shape[1]), self.image_size)
            if self._poll_interrupt is not None:
                self._poll_interrupt()
            latent = self.n.VAEEncode().encode(vae, px)[0]
            if self.image_size is not None:      # tile-local model patches (upscale/tile_ops.py:277)
                ctx = cropped_model_patches(model, (row.x1, row.y1, row.x2, row.y2), self.image_size)
            else:
                ctx = nullcontext(model)
            with ctx as tile_model:
                samples = self.n.common_ksampler(tile_model, seed, steps, cfg, sampler_name, scheduler, pos, neg,
                                                 latent, denoise=denoise)[0]
            if self.tiled_decode:
                img = self.n.VAEDecodeTiled().decode(vae, samples, tile_size=512)[0]
            else:
                img = self.n.VAEDecode().decode(vae, samples)[0]
            outs.append(img.to(tiles.device, torch.float32))
        return torch.stack(outs, 0)
