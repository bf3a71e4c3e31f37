"""Tile-local model patches for the real ComfyUI sampler.

Some ControlNet-style conditioners do not travel in the conditioning list but as *model patches*
(`model.model_options["transformer_options"]["patches"]`, classes `DiffSynthCnetPatch` /
`ZImageControlPatch`) that hold a full-canvas control image.  The reference cuts that image to
the tile's crop window around every sampler call and puts the patch back afterwards
(utils/crop_model_patch.py:10-114, used at upscale/tile_ops.py:277).  Same behaviour here, as a
context manager that restores the patch state deterministically on exit (the reference relies on
`__del__`).  Pure host logic on ComfyUI objects: slicing views, no pixel arithmetic.
"""
from __future__ import annotations

from contextlib import contextmanager
from typing import List, Sequence, Tuple

import torch

from .conditioning import Region, scale_region

PATCH_CLASSES = ("DiffSynthCnetPatch", "ZImageControlPatch")
_REQUIRED = ("image", "model_patch", "vae", "strength", "encoded_image", "encoded_image_size")


def _patch_objects(model) -> List[object]:
    """Unique croppable patch objects of a (cloned) model, in registration order."""
    options = getattr(model, "model_options", None) or {}
    table = options.get("transformer_options", {}).get("patches", {})
    seen, found = set(), []
    for group in table.values():
        for p in group:
            if id(p) in seen or type(p).__name__ not in PATCH_CLASSES:
                continue
            seen.add(id(p))
            found.append(p)
    return found


class _SavedPatch:
    def __init__(self, patch):
        lacking = [a for a in _REQUIRED if not hasattr(patch, a)]
        if lacking:
            raise AttributeError(f"{type(patch).__name__} missing required attrs: {', '.join(lacking)}")
        self.patch = patch
        keep = lambda v: v.clone() if isinstance(v, torch.Tensor) else v   # noqa: E731
        self.image, self.encoded, self.size = keep(patch.image), keep(patch.encoded_image), patch.encoded_image_size

    def restore(self):
        self.patch.image, self.patch.encoded_image, self.patch.encoded_image_size = self.image, self.encoded, self.size


def crop_patch(patch, regions: Sequence[Region], canvas_size: Tuple[int, int], latent_crop: bool = False):
    """Cut `patch.image` [B,H,W,C] (and, with latent_crop, `patch.encoded_image` [B,C,h,w]) to the
    regions (canvas coordinates), batch-concatenated in region order."""
    img = patch.image
    size = (img.shape[2], img.shape[1])                        # (W, H) of the control image
    boxes = [scale_region(r, canvas_size, size) for r in regions]
    cut = torch.cat([img[:, y1:y2, x1:x2, :] for (x1, y1, x2, y2) in boxes], dim=0)
    patch.image = cut
    patch.encoded_image_size = (cut.shape[1], cut.shape[2])
    if latent_crop:
        k = patch.vae.spacial_compression_encode()
        lat = patch.encoded_image
        patch.encoded_image = torch.cat([lat[:, :, y1 // k:y2 // k, x1 // k:x2 // k] for (x1, y1, x2, y2) in boxes], dim=0)
    else:   # let the patch re-encode the cropped image the way its constructor does
        patch.__init__(patch.model_patch, patch.vae, cut, patch.strength,
                       inpaint_image=getattr(patch, "inpaint_image", None), mask=getattr(patch, "mask", None))


@contextmanager
def cropped_model_patches(model, regions, canvas_size: Tuple[int, int], latent_crop: bool = False):
    """`with cropped_model_patches(model, (x1,y1,x2,y2), (W,H)) as m: common_ksampler(m, ...)`.
    Yields a clone whose croppable patches see only the tile; a model that cannot be cloned is
    yielded as is (utils/crop_model_patch.py:13-18); a patch that cannot be cropped is skipped
    (:35-39).  Patch state is restored on exit."""
    if regions and not isinstance(regions, list):
        regions = [tuple(regions)]
    try:
        clone = model.clone()
    except Exception:
        yield model
        return
    saved: List[_SavedPatch] = []
    for p in _patch_objects(clone):
        try:
            state = _SavedPatch(p)
        except Exception:
            continue
        try:
            crop_patch(p, regions, canvas_size, latent_crop)
            saved.append(state)
        except Exception:
            state.restore()
    try:
        yield clone
    finally:
        for s in reversed(saved):
            s.restore()
