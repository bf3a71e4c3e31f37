"""MODEL stand-ins that plug into UltimateSDUpscaleDistributed.run() when ComfyUI and its
checkpoints are not available (tests, smoke, benchmark)."""
from __future__ import annotations

from .denoise import SyntheticSDXL, T0Denoiser


class T0Model:
    """`model` argument whose sampler is the deterministic T0 denoiser."""

    def as_usdu_denoiser(self, seed, denoise, **_):
        return T0Denoiser(seed, denoise)


class SyntheticSDXLModel:
    def __init__(self, device="cuda", width=640, depth=6):
        self.device, self.width, self.depth = device, width, depth
        self._net = None

    def as_usdu_denoiser(self, steps, denoise, **_):
        if self._net is None:
            self._net = SyntheticSDXL(steps, self.width, self.depth, denoise).to(self.device, dtype=None)
            self._net = self._net.to(self.device).bfloat16()
        self._net.steps, self._net.denoise = steps, denoise
        return self._net
