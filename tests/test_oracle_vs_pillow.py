"""Pin the oracle's primitives bit-exactly against the installed Pillow (the third-party
library that owns the reference's pixel arithmetic)."""
import numpy as np
import pytest
from PIL import Image, ImageDraw, ImageFilter

import usdu_oracle as orc

RNG = np.random.default_rng(0)


@pytest.mark.parametrize("iw,ih,ow,oh", [
    (576, 576, 544, 544), (544, 544, 576, 576), (320, 320, 288, 288), (288, 288, 320, 320),
    (288, 288, 544, 544), (576, 300, 544, 544), (100, 37, 64, 64), (64, 64, 100, 37),
    (17, 9, 200, 3), (544, 544, 544, 600), (1, 1, 8, 8), (8, 8, 1, 1)])
def test_lanczos_matches_pillow(iw, ih, ow, oh):
    for kind in ("noise", "binary"):
        a = RNG.integers(0, 256, (ih, iw, 3), dtype=np.uint8)
        if kind == "binary":  # 0/255 patterns exercise clip8 overshoot
            a = (RNG.integers(0, 2, (ih, iw, 3)) * 255).astype(np.uint8)
        ref = np.array(Image.fromarray(a).resize((ow, oh), Image.LANCZOS))
        assert np.array_equal(orc.lanczos_resize_u8(a, ow, oh), ref)


@pytest.mark.parametrize("radius", [1, 2, 3, 4, 5, 7, 8, 15, 16, 31, 32, 64, 100, 255, 256])
def test_gaussian_blur_matches_pillow(radius):
    a = RNG.integers(0, 256, (24, 200), dtype=np.uint8)
    ref = np.array(Image.fromarray(a, "L").filter(ImageFilter.GaussianBlur(radius)))
    assert np.array_equal(orc.gaussian_blur_L(a, radius), ref)


def test_box_blur_params_every_integer_radius():
    a = np.zeros((3, 64), dtype=np.uint8)
    a[:, 20:40] = 255
    for r in range(1, 257):
        ref = np.array(Image.fromarray(a, "L").filter(ImageFilter.GaussianBlur(r)))
        assert np.array_equal(orc.gaussian_blur_L(a, r), ref), r


@pytest.mark.parametrize("W,H,x,y,tw,th,blur", [
    (700, 500, 256, 256, 256, 256, 8), (700, 500, 0, 0, 256, 256, 16), (700, 500, 512, 256, 256, 256, 32),
    (300, 200, 256, 128, 128, 128, 64), (300, 200, 128, 128, 128, 128, 0), (90, 70, 64, 64, 64, 64, 8),
    (90, 70, 0, 0, 128, 128, 8)])
def test_mask_window_equals_full_canvas_mask(W, H, x, y, tw, th, blur):
    m = Image.new("L", (W, H), 0)
    ImageDraw.Draw(m).rectangle([x, y, x + tw, y + th], fill=255)
    if blur > 0:
        m = m.filter(ImageFilter.GaussianBlur(blur))
    ref = np.array(m)
    assert np.array_equal(orc.tile_mask_full(W, H, x, y, tw, th, blur), ref)
    for win in [(0, 0, W, H), (max(x - 40, 0), max(y - 33, 0), min(x + tw + 40, W), min(y + th + 17, H))]:
        got = orc.tile_mask_window(W, H, x, y, tw, th, blur, win)
        assert np.array_equal(got, ref[win[1]:win[3], win[0]:win[2]])


def test_composite_matches_pillow_blend_sequence():
    H, W = 40, 50
    base = RNG.integers(0, 256, (H, W, 3), dtype=np.uint8)
    tile = RNG.integers(0, 256, (20, 30, 3), dtype=np.uint8)
    mask = RNG.integers(0, 256, (H, W), dtype=np.uint8)
    mask[:4] = 0
    mask[-4:] = 255
    x1, y1 = 7, 9
    # the reference's literal sequence, upscale/tile_ops.py:333-349
    layer = Image.new("RGBA", (W, H))
    layer.paste(Image.fromarray(tile), (x1, y1))
    tmp = layer.copy()
    tmp.putalpha(Image.fromarray(mask, "L"))
    layer.paste(tmp, layer)
    res = Image.fromarray(base).convert("RGBA")
    res.alpha_composite(layer)
    ref = np.array(res.convert("RGB"))
    got = base.copy()
    got[y1:y1 + 20, x1:x1 + 30] = orc.composite_u8(tile, base[y1:y1 + 20, x1:x1 + 30], mask[y1:y1 + 20, x1:x1 + 30])
    assert np.array_equal(got, ref)


def test_quantize_round_trip_all_codes():
    u = np.arange(256, dtype=np.uint8)
    assert np.array_equal(orc.quantize_u8(orc.dequantize_u8(u)), u)


@pytest.mark.parametrize("w,h,ow,oh", [(64, 48, 300, 200), (300, 200, 64, 48), (128, 128, 960, 540), (97, 31, 101, 77),
                                       (540, 960, 135, 240), (33, 33, 33, 70)])
def test_bicubic_resize_mode_L_matches_pillow(w, h, ow, oh):
    """Image.resize(..., BICUBIC) on mode L: what crop_mask applies to conditioning masks
    (utils/usdu_utils.py:424,435)."""
    rng = np.random.default_rng(w * 1000 + h)
    a = rng.integers(0, 256, (h, w), dtype=np.uint8)
    ref = np.array(Image.fromarray(a).resize((ow, oh), Image.Resampling.BICUBIC))
    assert np.array_equal(orc.resize_u8(a, ow, oh, "bicubic"), ref)
    ref = np.array(Image.fromarray(a).resize((ow, oh), Image.Resampling.LANCZOS))
    assert np.array_equal(orc.resize_u8(a, ow, oh, "lanczos"), ref)


@pytest.mark.parametrize("n_in,n_out", [(5, 17), (542, 576), (30, 7), (100, 100), (3, 1000), (1, 9), (254, 255), (255, 254)])
def test_nearest_index_matches_pillow(n_in, n_out):
    """Image.resize(..., NEAREST) along one axis (pad_image2's edge strips, utils/usdu_utils.py:190-199)."""
    a = (np.arange(n_in) % 251).astype(np.uint8)[None, :]
    ref = np.array(Image.fromarray(a).resize((n_out, 1), Image.Resampling.NEAREST))[0]
    assert np.array_equal(a[0][orc.nearest_index(n_in, n_out)], ref)
    ref = np.array(Image.fromarray(a.T.copy()).resize((1, n_out), Image.Resampling.NEAREST))[:, 0]
    assert np.array_equal(a[0][orc.nearest_index(n_in, n_out)], ref)
