"""Per-tile conditioning cropping == the reference's utils/usdu_utils.py functions, run side by
side on the same inputs (needs /root/reference; skipped on the GPU box)."""
import copy

import pytest
import torch

import ref_loader
from __graft_entry__ import load_package

load_package()
from comfyui_distributed_b200 import conditioning as C  # noqa: E402

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="reference tree not present")


class FakeControl:
    def __init__(self, hint, prev=None):
        self.cond_hint_original = hint
        self.previous_controlnet = prev

    def copy(self):
        return copy.copy(self)

    def set_previous_controlnet(self, p):
        self.previous_controlnet = p


def _ref():
    ref_loader.load()
    import sys
    return sys.modules[ref_loader.PKG + ".utils.usdu_utils"]


REGIONS = [((480, 992, 1056, 1568), (7680, 4320), (544, 544)), ((0, 0, 544, 544), (1300, 1100), (544, 544)),
           ((724, 524, 1300, 1100), (1300, 1100), (544, 544)), ((10, 20, 170, 150), (300, 260), (160, 136))]


@pytest.mark.parametrize("region,canvas,tile", REGIONS)
def test_control_hint_crop_matches_reference(region, canvas, tile):
    u = _ref()
    g = torch.Generator().manual_seed(1)
    h1 = torch.rand(1, 3, canvas[1] // 4, canvas[0] // 4, generator=g)
    h2 = torch.rand(2, 3, canvas[1] // 8 + 3, canvas[0] // 8 + 1, generator=g)
    mine = {"control": FakeControl(h1.clone(), FakeControl(h2.clone()))}
    theirs = {"control": FakeControl(h1.clone(), FakeControl(h2.clone()))}
    C.crop_control_hints(mine, region, canvas, tile)
    u.crop_controlnet(theirs, region, canvas, canvas, tile, 0, 0)
    a, b = mine["control"], theirs["control"]
    while b is not None:
        assert torch.equal(a.cond_hint_original, b.cond_hint_original)
        assert a.cond_hint_original.shape[-2:] == (tile[1], tile[0])
        a, b = a.previous_controlnet, b.previous_controlnet
    assert a is None


@pytest.mark.parametrize("region,canvas,tile", REGIONS)
def test_area_gligen_reflatents_match_reference(region, canvas, tile):
    u = _ref()
    init = (canvas[0] // 2, canvas[1] // 2)
    for area in [(40, 60, 10, 20), (8, 8, 0, 0), (500, 500, 3, 7), (1, 1, 400, 400)]:
        mine, theirs = {"area": area, "strength": 1.0}, {"area": area, "strength": 1.0}
        C.crop_area(mine, region, init, canvas, 0, 0)
        u.crop_area(theirs, region, init, canvas, tile, 0, 0)
        assert mine == theirs
    boxes = [("e1", 20, 30, 5, 6), ("e2", 64, 64, 60, 90), ("e3", 4, 4, 500, 500)]
    mine, theirs = {"gligen": ("position", "m", list(boxes))}, {"gligen": ("position", "m", list(boxes))}
    C.crop_gligen(mine, region, init, canvas, 0, 0)
    u.crop_gligen(theirs, region, init, canvas, tile, 0, 0)
    assert mine == theirs
    g = torch.Generator().manual_seed(2)
    lat = [torch.rand(1, 4, canvas[1] // 8, canvas[0] // 8, generator=g), torch.rand(1, 4, 1, 40, 50, generator=g)]
    mine, theirs = {"reference_latents": [t.clone() for t in lat]}, {"reference_latents": [t.clone() for t in lat]}
    C.crop_reference_latents(mine, region, canvas, tile)
    u.crop_reference_latents(theirs, region, init, canvas, tile, 0, 0)
    for a, b in zip(mine["reference_latents"], theirs["reference_latents"]):
        assert torch.equal(a, b)


def test_crop_cond_and_clone_do_not_touch_the_originals():
    hint = torch.rand(1, 3, 64, 64)
    cond = [[torch.rand(1, 77, 8), {"control": FakeControl(hint), "area": (8, 8, 0, 0), "pooled_output": torch.rand(1, 8)}]]
    keep = hint.clone()
    out = C.crop_cond(C.clone_conditioning(cond), (0, 0, 128, 128), (256, 256), (256, 256), (128, 128))
    assert torch.equal(hint, keep) and cond[0][1]["control"].cond_hint_original is hint
    assert out[0][1]["control"].cond_hint_original.shape == (1, 3, 128, 128)
    if not torch.cuda.is_available():        # masks are cropped on the GPU only: loud failure, never a CPU path
        from comfyui_distributed_b200._native import NativeError
        with pytest.raises(NativeError):
            C.crop_cond([[None, {"mask": torch.rand(1, 8, 8)}]], (0, 0, 8, 8), (8, 8), (8, 8), (8, 8))
