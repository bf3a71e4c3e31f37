"""Per-tile conditioning for the real ComfyUI sampler.

Text embeddings are shared between tiles; entries that carry SPATIAL hints must be cut to the
tile's crop window before sampling, as the reference does for every tile
(upscale/conditioning.py:17-34 clone, utils/usdu_utils.py:297-312 ControlNet hints, :335-378
GLIGEN boxes, :381-412 areas, :445-503 reference latents, :506-517 crop_cond).  Everything
here is torch / integer arithmetic on whatever device the hints live on -- no PIL.  Mask
conditioning (:415-442, PIL BICUBIC + edge padding) is not ported yet and raises instead of
sampling with an uncropped mask.
"""
from __future__ import annotations

import copy
import math
from typing import Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Region = Tuple[int, int, int, int]


def scale_region(region: Region, from_size: Sequence[int], to_size: Sequence[int]) -> Region:
    """Map a rectangle from an image of `from_size` (w, h) to the same image at `to_size`:
    floor the near edges, ceil the far ones (utils/usdu_utils.py:115-124)."""
    x1, y1, x2, y2 = region
    fw, fh = from_size
    tw, th = to_size
    return (math.floor(x1 * tw / fw), math.floor(y1 * th / fh), math.ceil(x2 * tw / fw), math.ceil(y2 * th / fh))


def intersect(a: Region, b: Region) -> Optional[Region]:
    x1, y1, x2, y2 = max(a[0], b[0]), max(a[1], b[1]), min(a[2], b[2]), min(a[3], b[3])
    return None if (x1 >= x2 or y1 >= y2) else (x1, y1, x2, y2)


def _clone_control_chain(control, clone_hint: bool):
    if control is None:
        return None
    new = copy.copy(control)
    if clone_hint and getattr(control, "cond_hint_original", None) is not None:
        new.cond_hint_original = control.cond_hint_original.clone()
    if hasattr(control, "previous_controlnet"):
        new.previous_controlnet = _clone_control_chain(control.previous_controlnet, clone_hint)
    return new


def clone_conditioning(cond, clone_hints: bool = True):
    """New list / dicts / hint tensors per tile, models shared (upscale/conditioning.py:17-34)."""
    out = []
    for emb, opts in cond:
        d = dict(opts)
        if "control" in d:
            d["control"] = _clone_control_chain(d["control"], clone_hints)
        for key in ("mask", "pooled_output"):
            if d.get(key) is not None:
                d[key] = d[key].clone()
        if "area" in d:
            d["area"] = d["area"][:]
        out.append([emb.clone() if emb is not None else None, d])
    return out


def crop_control_hints(opts: dict, region: Region, canvas_size, tile_size):
    """Every ControlNet of the chain gets its hint [B,C,h,w] cut to the window (scaled to the
    hint's own resolution) and resized nearest-exact to the tile size."""
    c = opts.get("control")
    if c is None:
        return
    head = copy.copy(c)
    opts["control"] = head
    node = head
    while node is not None:
        hint = node.cond_hint_original
        hx1, hy1, hx2, hy2 = scale_region(region, canvas_size, (hint.shape[-1], hint.shape[-2]))
        hint = hint[:, :, hy1:hy2, hx1:hx2]
        node.cond_hint_original = F.interpolate(hint, size=(tile_size[1], tile_size[0]), mode="nearest-exact")
        prev = getattr(node, "previous_controlnet", None)
        prev = copy.copy(prev) if prev is not None else None
        if hasattr(node, "set_previous_controlnet"):
            node.set_previous_controlnet(prev)
        else:
            node.previous_controlnet = prev
        node = prev


def _to_tile_box(box: Region, region: Region, w_pad: int, h_pad: int) -> Region:
    return (box[0] - region[0] + w_pad, box[1] - region[1] + h_pad, box[2] - region[0] + w_pad, box[3] - region[1] + h_pad)


def crop_gligen(opts: dict, region: Region, init_size, canvas_size, w_pad: int = 0, h_pad: int = 0):
    if "gligen" not in opts:
        return
    kind, model, boxes = opts["gligen"]
    if kind != "position":
        return
    kept = []
    for emb, h, w, y, x in boxes:
        box = scale_region((x * 8, y * 8, (x + w) * 8, (y + h) * 8), init_size, canvas_size)
        hit = intersect(box, region)
        if hit is None:
            continue
        x1, y1, x2, y2 = _to_tile_box(hit, region, w_pad, h_pad)
        kept.append((emb, (y2 - y1) // 8, (x2 - x1) // 8, y1 // 8, x1 // 8))
    opts["gligen"] = (kind, model, kept)


def crop_area(opts: dict, region: Region, init_size, canvas_size, w_pad: int = 0, h_pad: int = 0):
    if "area" not in opts:
        return
    h, w, y, x = opts["area"]
    box = scale_region((8 * x, 8 * y, 8 * (x + w), 8 * (y + h)), init_size, canvas_size)
    hit = intersect(box, region)
    if hit is None:
        del opts["area"]
        opts.pop("strength", None)
        return
    x1, y1, x2, y2 = _to_tile_box(hit, region, w_pad, h_pad)
    opts["area"] = ((y2 - y1) // 8, (x2 - x1) // 8, y1 // 8, x1 // 8)


def crop_reference_latents(opts: dict, region: Region, canvas_size, tile_size, k: int = 8):
    lat = opts.get("reference_latents")
    if not isinstance(lat, list):
        return
    cw, chh = canvas_size[0] // k, canvas_size[1] // k
    tw, th = max(1, tile_size[0] // k), max(1, tile_size[1] // k)
    x1, y1, x2, y2 = region
    out = []
    for t in lat:
        five = t.ndim == 5
        if five:
            t = t.squeeze(2)
        if t.ndim != 4:
            raise ValueError(f"expected BCHW or BC1HW, got {tuple(t.shape)}")
        if tuple(t.shape[-2:]) != (chh, cw):
            t = F.interpolate(t, size=(chh, cw), mode="bilinear", align_corners=False)
        t = t[:, :, int(round(y1 / k)):int(round(y2 / k)), int(round(x1 / k)):int(round(x2 / k))]
        t = F.interpolate(t, size=(th, tw), mode="bilinear", align_corners=False)
        out.append(t.unsqueeze(2) if five else t)
    opts["reference_latents"] = out


def crop_cond(cond, region: Region, init_size, canvas_size, tile_size, w_pad: int = 0, h_pad: int = 0):
    """utils/usdu_utils.py:506-517 on an already cloned conditioning list."""
    out = []
    for emb, opts in cond:
        d = dict(opts)
        if d.get("mask") is not None:
            raise NotImplementedError("mask conditioning needs per-tile BICUBIC cropping (utils/usdu_utils.py:415-442): not ported yet")
        crop_control_hints(d, region, canvas_size, tile_size)
        crop_gligen(d, region, init_size, canvas_size, w_pad, h_pad)
        crop_area(d, region, init_size, canvas_size, w_pad, h_pad)
        crop_reference_latents(d, region, canvas_size, tile_size)
        out.append([emb, d])
    return out


def make_cond_cropper():
    """-> fn(positive, negative, tile, tile_size (w,h), image_size (w,h)) used by ComfySampler,
    mirroring process_tiles_batch (upscale/tile_ops.py:263-273)."""
    def crop(positive, negative, tile, tile_size, image_size):
        region = (tile.x1, tile.y1, tile.x2, tile.y2)
        pos = crop_cond(clone_conditioning(positive), region, image_size, image_size, tile_size)
        neg = crop_cond(clone_conditioning(negative), region, image_size, image_size, tile_size)
        return pos, neg
    return crop
