"""GPU parity of the conditioning-mask crop (csrc/usdu_plane.cu through conditioning.MaskCropper)
against the fixtures made by the reference's crop_mask (utils/usdu_utils.py:415-442) and against the
oracle on sizes the fixtures do not cover (8K canvas window).  Bit-exact."""
import os

import numpy as np
import pytest
import torch

import usdu_oracle as orc
from __graft_entry__ import load_package
from inputs import MASK_CROP_CASES, make_mask

load_package()
from comfyui_distributed_b200 import _native as nat  # noqa: E402
from comfyui_distributed_b200 import conditioning as C  # noqa: E402
from comfyui_distributed_b200.planner import get_plan  # noqa: E402

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")
DEV = "cuda:0"


def _u8(x: torch.Tensor) -> np.ndarray:
    v = x.cpu().numpy()
    q = np.round(v * 255).astype(np.uint8)
    assert np.array_equal(q.astype(np.float32) / np.float32(255), v)      # values are exactly k/255
    return q


@pytest.mark.parametrize("case", MASK_CROP_CASES, ids=lambda c: c[0])
@pytest.mark.parametrize("where", ["cpu", "cuda"])
def test_mask_crop_matches_reference_fixture(case, where):
    name, kind, seed, B, (Hm, Wm), region, canvas, tile = case
    gold = np.load(os.path.join(G, "mask_crop.npz"))[name]
    m = torch.from_numpy(make_mask(kind, seed, B, Hm, Wm)).to(where)
    before = m.clone()
    out = C.MaskCropper(DEV).crop(m, region, canvas, tile)
    assert out.is_cuda and out.dtype == torch.float32 and tuple(out.shape) == (B, tile[1], tile[0])
    assert np.array_equal(_u8(out), gold)
    assert torch.equal(m, before)                                          # caller's tensor untouched


def test_mask_crop_every_tile_of_an_8k_plan():
    """cfg2 geometry: a 540x960 mask, every 9th tile of the 135 (interior, edges, corners)."""
    W, H = 7680, 4320
    plan = get_plan(W, H, 512, 512, 32, 8, True)
    m = make_mask("blob", 11, 1, 540, 960)
    q = orc.quantize_u8(m[0])
    cropper = C.MaskCropper(DEV)
    mt = torch.from_numpy(m)
    for t in plan.tiles[::9] + [plan.tiles[-1]]:
        region = (t.x1, t.y1, t.x2, t.y2)
        got = _u8(cropper.crop(mt, region, (W, H), (t.pw, t.ph)))[0]
        # the oracle resizes only what the window needs too (full 8K BICUBIC per tile would take minutes):
        want = _window_oracle(q, region, (W, H), (t.pw, t.ph))
        assert np.array_equal(got, want), t.id
    assert len(cropper._masks) == 1                                         # quantised once, reused


def _window_oracle(q, region, canvas, tile):
    """oracle.crop_mask_u8 with the BICUBIC upscale evaluated for the window only (same taps)."""
    W, H = canvas
    x1, y1, x2, y2 = region
    Hm, Wm = q.shape
    bh, kh = orc.resample_coeffs(Wm, W, "bicubic")
    bv, kv = orc.resample_coeffs(Hm, H, "bicubic")
    def axis(src, bounds, kk, lo, hi):      # resample along axis 0, outputs lo..hi
        out = np.zeros((hi - lo,) + src.shape[1:], dtype=np.uint8)
        for o in range(lo, hi):
            a, n = bounds[o]
            acc = (src[a:a + n].astype(np.int64) * kk[o, :n].astype(np.int64).reshape((n,) + (1,) * (src.ndim - 1))).sum(0) + (1 << 21)
            out[o - lo] = np.clip(acc >> 22, 0, 255)
        return out
    mid = axis(q.T, bh, kh, x1, x2).T if Wm != W else q[:, x1:x2]
    win = axis(mid, bv, kv, y1, y2) if Hm != H else mid[y1:y2]
    pw, ph = tile
    rw, rh, hp, vp = orc.mask_fit_geometry(x2 - x1, y2 - y1, pw, ph)
    m = orc.resize_u8(win, rw, rh, "lanczos")
    m = orc.pad_fill_u8(m, hp, vp)
    return orc.resize_u8(m, pw, ph, "lanczos")


def test_crop_cond_with_mask_and_cropper_cache():
    W, H = 1300, 1100
    plan = get_plan(W, H, 512, 512, 32, 8, True)
    m = torch.from_numpy(make_mask("noise", 3, 2, 110, 130))
    cond = [[torch.zeros(1, 77, 8), {"mask": m, "pooled_output": torch.zeros(1, 8)}]]
    crop = C.make_cond_cropper()
    for t in plan.tiles[:3]:
        pos, neg = crop(cond, cond, t, (t.pw, t.ph), (W, H))
        want = np.stack([orc.crop_mask_u8(orc.quantize_u8(m[b].numpy()), (t.x1, t.y1, t.x2, t.y2), (W, H), (t.pw, t.ph))
                         for b in range(2)])
        assert np.array_equal(_u8(pos[0][1]["mask"]), want)
        assert np.array_equal(_u8(neg[0][1]["mask"]), want)
    assert cond[0][1]["mask"] is m                                         # the caller's conditioning is not edited


def test_plane_kernels_reject_bad_arguments():
    with pytest.raises(nat.NativeError):
        nat.plane_resample_u8(None, 1, 8, 8, 8, 64, None, 0, 8, None, 0, 8, 0, 0, None, None, 8, 64, 0)
    x = torch.zeros((1, 8, 8), dtype=torch.uint8, device=DEV)
    with pytest.raises(nat.NativeError):     # window outside the source when no horizontal table is given
        nat.plane_resample_u8(x.data_ptr(), 1, 8, 8, 8, 64, None, 4, 8, None, 0, 8, 0, 0, None, x.data_ptr(), 8, 64, 0)
    with pytest.raises(nat.NativeError):     # side pads without an index table
        nat.plane_pad_fill_u8(x.data_ptr(), 1, 8, 8, 8, 64, 2, 0, None, None, x.data_ptr(), 12, 96, 0)
