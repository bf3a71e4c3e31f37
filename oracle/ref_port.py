"""Cost-faithful CPU port of the reference's tile path -- TEST / BASELINE INFRASTRUCTURE ONLY.

``usdu_oracle.py`` restates the reference's ARITHMETIC window-by-window (fast, numpy).
This file restates its COST STRUCTURE: the same sequence of Pillow calls on the same
full-canvas-sized temporaries, once per tile, that the reference executes on the host
(SURVEY.md section 3.2 / 6).  It is what ``bench.py`` times as ``cpu_baseline`` and as
the ``--impl reference`` arm on the GPU box, where /root/reference does not exist (the
reference is pure Python and cannot be "compiled into oracle/_ref").  Outputs are
bit-identical to the real reference (tests/test_ref_port.py checks it against the
fixtures generated from the real code).

Call-site map (reference @ a91f9fb):
  to_pil / to_tensor ............ utils/image.py:8-18
  full-canvas mask + getbbox .... upscale/tile_ops.py:108-112, utils/usdu_utils.py:49-62
  LANCZOS resize ................ upscale/tile_ops.py:141-150
  create_tile_mask .............. upscale/tile_ops.py:289-308
  blend_tile .................... upscale/tile_ops.py:310-349
  progressive driver ............ upscale/modes/single_gpu.py:8-72
  static worker / master loops .. upscale/modes/static.py:191-314, :371-570
  PNG level-0 tile payloads ..... upscale/worker_comms.py:25-46, upscale/payload_parsers.py:7-64
"""
from __future__ import annotations

import io
import time
from typing import Callable, Dict, List, Optional, Sequence

import numpy as np
import torch
from PIL import Image, ImageDraw, ImageFilter

import usdu_oracle as orc


def to_pil(batch: torch.Tensor, index: int = 0) -> Image.Image:
    return Image.fromarray((255 * batch[index].cpu().numpy()).astype(np.uint8))


def to_tensor(img: Image.Image) -> torch.Tensor:
    return torch.from_numpy(np.array(img).astype(np.float32) / 255.0).unsqueeze(0)


def torch_t0(seed: int, denoise: float) -> Callable[[torch.Tensor], torch.Tensor]:
    """The T0 denoiser on CPU torch tensors (same arithmetic as orc.make_t0_denoiser)."""
    d = float(np.float32(denoise))
    omd = float(np.float32(1.0) - np.float32(denoise))
    cache: Dict[tuple, torch.Tensor] = {}

    def fn(px: torch.Tensor) -> torch.Tensor:
        key = tuple(px.shape)
        if key not in cache:
            cache[key] = torch.from_numpy(orc.t0_noise(seed, key)) * d
        return torch.clamp(px * omd + cache[key], 0.0, 1.0)

    return fn


class Timer:
    def __init__(self):
        self.t: Dict[str, float] = {}

    def add(self, key: str, dt: float):
        self.t[key] = self.t.get(key, 0.0) + dt


class RefPort:
    """One participant's view of the job (canvas list, masks, tile grid)."""

    def __init__(self, W: int, H: int, tile_width: int, tile_height: int, padding: int, mask_blur: int,
                 uniform: bool):
        self.W, self.H, self.padding, self.blur, self.uniform = W, H, padding, mask_blur, uniform
        self.tw, self.th = orc.round_to_multiple(tile_width), orc.round_to_multiple(tile_height)
        self.grid = orc.calculate_tiles(W, H, self.tw, self.th)
        self.timer = Timer()

    # -- the reference builds a canvas-sized 'L' image per tile just to get a bbox -----
    def _window(self, x: int, y: int):
        t0 = time.perf_counter()
        probe = Image.new("L", (self.W, self.H), 0)
        ImageDraw.Draw(probe).rectangle([x, y, x + self.tw, y + self.th], fill=255)
        box = probe.getbbox()
        self.timer.add("bbox", time.perf_counter() - t0)
        assert box == orc._rect_bbox(self.W, self.H, x, y, self.tw, self.th)
        return orc.crop_geometry(self.W, self.H, x, y, self.tw, self.th, self.padding, self.uniform)

    def feather(self, x: int, y: int) -> Image.Image:
        t0 = time.perf_counter()
        m = Image.new("L", (self.W, self.H), 0)
        ImageDraw.Draw(m).rectangle([x, y, x + self.tw, y + self.th], fill=255)
        if self.blur > 0:
            m = m.filter(ImageFilter.GaussianBlur(self.blur))
        self.timer.add("mask", time.perf_counter() - t0)
        return m

    def extract(self, frames: List[Image.Image], x: int, y: int):
        """Full-canvas fp32 round trip of every frame (single_gpu.py:42), then crop+resize."""
        t0 = time.perf_counter()
        source = torch.cat([to_tensor(f) for f in frames], dim=0)
        self.timer.add("canvas_to_tensor", time.perf_counter() - t0)
        x1, y1, x2, y2, pw, ph = self._window(x, y)
        t0 = time.perf_counter()
        crop = source[:, y1:y2, x1:x2, :]
        out = []
        for b in range(crop.shape[0]):
            im = to_pil(crop, b)
            if im.size != (pw, ph):
                im = im.resize((pw, ph), Image.LANCZOS)
            out.append(to_tensor(im))
        self.timer.add("crop_resize", time.perf_counter() - t0)
        return torch.cat(out, dim=0), (x1, y1, x2 - x1, y2 - y1)

    def blend(self, base: Image.Image, tile: Image.Image, x1: int, y1: int, ew: int, eh: int,
              mask: Image.Image) -> Image.Image:
        t0 = time.perf_counter()
        if tile.size != (ew, eh):
            tile = tile.resize((ew, eh), Image.LANCZOS)
        layer = Image.new("RGBA", base.size)
        layer.paste(tile, (x1, y1))
        with_alpha = layer.copy()
        with_alpha.putalpha(mask)
        layer.paste(with_alpha, layer)
        out = base.convert("RGBA")
        out.alpha_composite(layer)
        out = out.convert("RGB")
        self.timer.add("blend", time.perf_counter() - t0)
        return out

    # -- drivers -----------------------------------------------------------------------
    def run_tiles(self, frames: List[Image.Image], tile_ids: Sequence[int], masks: Dict[int, Image.Image],
                  denoise: Callable[[torch.Tensor], torch.Tensor], keep: Optional[dict] = None):
        """Progressive loop over `tile_ids` on this participant's own canvas."""
        for tid in tile_ids:
            x, y = self.grid[tid]
            batch, (x1, y1, ew, eh) = self.extract(frames, x, y)
            t0 = time.perf_counter()
            done = denoise(batch)
            self.timer.add("denoise", time.perf_counter() - t0)
            for b in range(len(frames)):
                frames[b] = self.blend(frames[b], to_pil(done, b), x1, y1, ew, eh, masks[tid])
            if keep is not None:
                keep[tid] = (done, x1, y1, ew, eh)
        return frames


def process_single(image: torch.Tensor, denoise, tile_width, tile_height, padding, mask_blur, uniform=True,
                   max_tiles: Optional[int] = None, time_budget_s: Optional[float] = None,
                   timer_out: Optional[dict] = None) -> torch.Tensor:
    """process_single_gpu with the reference's cost structure.  ``max_tiles`` /
    ``time_budget_s`` bound the run to the first tiles of the grid for the benchmark's
    bounded sample (the reference builds all masks up front, single_gpu.py:35-37; here a
    tile's mask is built right before the tile so a truncated run pays only for the tiles
    it processes -- the per-tile total is the same)."""
    B, H, W, _ = image.shape
    port = RefPort(W, H, tile_width, tile_height, padding, mask_blur, uniform)
    ids = list(range(len(port.grid)))
    if max_tiles is not None:
        ids = ids[:max_tiles]
    t0 = time.perf_counter()
    frames = [to_pil(image[b:b + 1], 0).convert("RGB").copy() for b in range(B)]
    port.timer.add("q0", time.perf_counter() - t0)
    start = time.perf_counter()
    done_ids = []
    for tid in ids:
        masks = {tid: port.feather(*port.grid[tid])}
        frames = port.run_tiles(frames, [tid], masks, denoise)
        done_ids.append(tid)
        if time_budget_s is not None and time.perf_counter() - start > time_budget_s:
            break
    ids = done_ids
    t0 = time.perf_counter()
    res = torch.cat([to_tensor(f) for f in frames], dim=0)
    port.timer.add("result", time.perf_counter() - t0)
    if timer_out is not None:
        timer_out.update(port.timer.t)
        timer_out["tiles_done"] = len(ids)
        timer_out["tiles_total"] = len(port.grid)
    return res


# -- payload codec of the HTTP path (PNG, compress_level 0) -----------------------------
def encode_tile_png(tile: torch.Tensor) -> bytes:
    buf = io.BytesIO()
    to_pil(tile, 0).save(buf, format="PNG", compress_level=0)
    return buf.getvalue()


def decode_tile_png(data: bytes) -> Image.Image:
    return Image.open(io.BytesIO(data)).convert("RGB")
