"""One-GPU tests of the entry points the multi-GPU job is built from (so the driver's single-GPU run exercises them):
row-range quantise / dequantise, the one-launch slab gathers (here the "peers" are three canvases on the same device),
and the crop straight from the fp32 image."""
import numpy as np
import pytest
import torch

import usdu_oracle as orc
from __graft_entry__ import load_package
from inputs import make_input

load_package()
from comfyui_distributed_b200 import _native as nat  # noqa: E402
from comfyui_distributed_b200 import engine, planner  # noqa: E402

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _stream():
    return torch.cuda.current_stream().cuda_stream


@pytest.mark.parametrize("B,H,W,y0,y1", [(1, 64, 64, 8, 40), (2, 33, 50, 0, 33), (3, 40, 1021, 17, 18), (1, 540, 960, 500, 540), (2, 16, 8, 5, 5)])
def test_quantize_and_dequantize_rows_touch_only_their_rows(B, H, W, y0, y1):
    rng = np.random.default_rng(1)
    img = rng.random((B, H, W, 3), dtype=np.float32)
    x = torch.from_numpy(img).to(DEV)
    pitch = (W * 3 + 127) // 128 * 128
    canvas = torch.full((B, H, pitch), 7, dtype=torch.uint8, device=DEV)
    nat.quantize_rows(x.data_ptr(), canvas.data_ptr(), B, H, W, pitch, y0, y1, _stream())
    got = canvas[:, :, :W * 3].reshape(B, H, W, 3).cpu().numpy()
    want = np.full_like(got, 7)
    want[:, y0:y1] = orc.quantize_u8(img)[:, y0:y1]
    assert np.array_equal(got, want)
    back = torch.full((B, H, W, 3), -1.0, dtype=torch.float32, device=DEV)
    nat.dequantize_rows(canvas.data_ptr(), back.data_ptr(), B, H, W, pitch, y0, y1, _stream())
    wantf = np.full((B, H, W, 3), -1.0, dtype=np.float32)
    wantf[:, y0:y1] = orc.dequantize_u8(want[:, y0:y1])
    assert np.array_equal(back.cpu().numpy(), wantf)


@pytest.mark.parametrize("B,H,W,cuts", [(1, 96, 256, (0, 32, 64, 96)), (2, 50, 300, (0, 1, 49, 50)), (1, 40, 1021, (0, 10, 40)),
                                        (1, 37, 85, (0, 37)), (3, 24, 64, (0, 8, 8, 24))])
def test_gather_dequantize_and_gather_canvas_take_each_slab_from_its_owner(B, H, W, cuts):
    """rows [cuts[q], cuts[q+1]) come from canvas q (on a multi-GPU box: a peer's HBM over NVLink); widths that are not
    multiples of 4 go through the scalar fallback; empty slabs are allowed."""
    rng = np.random.default_rng(2)
    pitch = (W * 3 + 127) // 128 * 128
    n = len(cuts) - 1
    canv = [torch.from_numpy(rng.integers(0, 256, (B, H, pitch), dtype=np.uint8)).to(DEV) for _ in range(n)]
    want_u8 = np.zeros((B, H, W * 3), dtype=np.uint8)
    for q in range(n):
        want_u8[:, cuts[q]:cuts[q + 1]] = canv[q][:, cuts[q]:cuts[q + 1], :W * 3].cpu().numpy()
    out = torch.full((B, H, W, 3), -1.0, dtype=torch.float32, device=DEV)
    nat.gather_dequantize([c.data_ptr() for c in canv], list(cuts), out.data_ptr(), B, H, W, pitch, _stream())
    assert np.array_equal(out.cpu().numpy(), orc.dequantize_u8(want_u8).reshape(B, H, W, 3))
    mine = canv[0].clone()                                      # all-gather into canvas 0: its own slab stays, the others arrive
    nat.gather_canvas([mine.data_ptr()] + [c.data_ptr() for c in canv[1:]], list(cuts), mine.data_ptr(), B, H, W, pitch, _stream())
    assert np.array_equal(mine[:, :, :W * 3].cpu().numpy(), want_u8)


@pytest.mark.parametrize("kind,B,H,W,tile,pad", [("noise", 1, 512, 512, 256, 32), ("noise", 2, 260, 320, 128, 16), ("smooth", 1, 1100, 1304, 512, 32)])
def test_crop_from_the_fp32_image_equals_crop_from_the_quantised_canvas(kind, B, H, W, tile, pad):
    """usdu_tile_crop_resize_f32 (the truncating cast happens while the window is staged) against the oracle and against the
    crop of the quantised canvas, bit for bit -- what lets a conflict-free rank skip the whole-canvas quantise."""
    p = planner.Plan.build(W, H, tile, tile, pad, 8, True)
    dp = engine.DevicePlan.get(p, torch.device(DEV))
    img = make_input(kind, 5, B, H, W)
    x = torch.from_numpy(img).to(DEV)
    canvas = engine.Canvas(dp, B).load(x)
    if not canvas.can_crop_image():
        pytest.skip("no tensor-core crop for this plan")
    ids = list(range(len(p.tiles)))
    a, offs = canvas.crop(ids)
    b, offs_b = canvas.crop(ids, image=x)
    assert np.array_equal(offs, offs_b) and torch.equal(a, b)
    cu8 = orc.quantize_u8(img)
    _, _, oplan = orc.make_plan(W, H, tile, tile, pad, True)
    for t, o in zip(oplan, offs):
        want = orc.extract_tile(cu8, t)
        assert np.array_equal(b[int(o): int(o) + want.size].cpu().numpy().reshape(want.shape), want), t.idx
