"""Synthetic canvases shared by the golden generator and the tests (values k/255 fp32)."""
import numpy as np
import torch


def make_input(kind: str, seed: int, B: int, H: int, W: int) -> np.ndarray:
    if kind == "noise":
        g = torch.Generator().manual_seed(seed)
        x = torch.rand(B, H, W, 3, generator=g)
        return (torch.floor(x * 255) / 255).numpy().astype(np.float32)
    if kind == "smooth":
        g = torch.Generator().manual_seed(seed)
        yy, xx = np.mgrid[0:H, 0:W].astype(np.float32)
        base = np.stack([xx / max(W - 1, 1), yy / max(H - 1, 1), (xx + yy) / max(W + H - 2, 1)], -1)
        n = torch.rand(B, H, W, 3, generator=g).numpy() * 0.1
        v = np.clip(base[None] * 0.9 + n, 0, 1)
        return (np.floor(v * 255) / 255).astype(np.float32)
    if kind == "checker":
        yy, xx = np.mgrid[0:H, 0:W]
        c = (((xx // 3) + (yy // 5)) % 2).astype(np.float32)
        img = np.stack([c, 1 - c, c], -1)[None].repeat(B, 0)
        return img.astype(np.float32)
    raise ValueError(kind)
