"""Per-tile conditioning for the real ComfyUI sampler (upscale/conditioning.py:17-34,
utils/usdu_utils.py:506-517).  Text embeddings are shared; entries that carry spatial
hints (ControlNet hints, masks, areas, GLIGEN boxes, reference latents) need cropping to
the tile window.  Round 1 handles the hint-free case exactly (a per-tile shallow copy)
and raises for spatial hints instead of silently sampling with uncropped ones."""
from __future__ import annotations

_SPATIAL_KEYS = ("control", "gligen", "area", "mask", "reference_latents")


def clone_conditioning(cond):
    out = []
    for emb, opts in cond:
        d = dict(opts)
        if d.get("pooled_output") is not None:
            d["pooled_output"] = d["pooled_output"].clone()
        out.append([emb.clone() if emb is not None else None, d])
    return out


def make_cond_cropper():
    def crop(positive, negative, tile, tile_size, image_size):
        for cond in (positive, negative):
            for _, opts in cond:
                hit = [k for k in _SPATIAL_KEYS if opts.get(k) is not None]
                if hit:
                    raise NotImplementedError(
                        f"conditioning carries spatial hints {hit}: per-tile hint cropping "
                        "(utils/usdu_utils.py:297-503) is not implemented yet")
        return clone_conditioning(positive), clone_conditioning(negative)
    return crop
