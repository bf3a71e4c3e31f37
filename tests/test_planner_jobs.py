"""The planner's fast-path job records, interpreted by a numpy model of the kernels (tests/kernel_model.py),
reproduce the oracle: crop + LANCZOS of every tile, and the ordered seam blend (single launch with
overlapping tiles, per-wave launches, partitioned launches).  Checks on a machine WITHOUT a GPU what the
device tests check with one: staging windows, table rows, clip boxes, mask offsets, chain order, addresses."""
import numpy as np
import pytest

import kernel_model as km
import usdu_oracle as orc
from __graft_entry__ import load_package
from inputs import make_input

load_package()
from comfyui_distributed_b200 import planner  # noqa: E402

CASES = [("noise", 1, 300, 420, 128, 16, 8, True), ("smooth", 2, 260, 300, 128, 32, 16, True),
         ("noise", 1, 200, 232, 96, 16, 8, False), ("checker", 1, 150, 530, 64, 8, 4, True),
         ("noise", 1, 333, 257, 128, 0, 8, True), ("noise", 1, 96, 100, 128, 32, 8, True)]


def _ids(c):
    return f"{c[0]}-b{c[1]}-{c[3]}x{c[2]}-t{c[4]}-p{c[5]}-m{c[6]}-{'u' if c[7] else 'n'}"


@pytest.mark.parametrize("case", CASES, ids=_ids)
def test_crop_records_reproduce_extract_tile(case):
    kind, B, H, W, tile, pad, blur, uniform = case
    p = planner.Plan.build(W, H, tile, tile, pad, blur, uniform)
    if not p.fast:
        pytest.skip("this geometry runs on the generic kernels")
    canvas = orc.quantize_u8(make_input(kind, 3, B, H, W))
    _, _, oplan = orc.make_plan(W, H, tile, tile, pad, uniform)
    ids = list(range(len(p.tiles)))
    wl, offs, total = p.crop_worklist(ids, B, True)
    out = np.full(total, -1.0, dtype=np.float32)
    km.run_crop(p, canvas, wl, out)
    for t, o in zip(oplan, offs):
        want = orc.extract_tile(canvas, t)
        got = out[o:o + want.size].reshape(want.shape)
        assert np.array_equal(got, want), t.idx
    assert not (out < 0).any()                       # every element of every tile slot was written exactly by some block


@pytest.mark.parametrize("src_u8", [False, True])
@pytest.mark.parametrize("case", CASES, ids=_ids)
def test_blend_records_reproduce_ordered_blend(case, src_u8):
    kind, B, H, W, tile, pad, blur, uniform = case
    p = planner.Plan.build(W, H, tile, tile, pad, blur, uniform)
    if not p.fast:
        pytest.skip("this geometry runs on the generic kernels")
    tw, th, oplan = orc.make_plan(W, H, tile, tile, pad, uniform)
    base = orc.quantize_u8(make_input(kind, 4, B, H, W))
    rng = np.random.default_rng(1)
    ids = list(range(len(p.tiles)))
    offs, total = p.slot_offsets(ids, B)
    src = rng.random(total, dtype=np.float32)
    pool = km.mask_pool(p)
    want = base.copy()
    for t, o in zip(oplan, offs):
        proc = src[o:o + B * t.ph * t.pw * 3].reshape(B, t.ph, t.pw, 3)
        m = orc.tile_mask_window(W, H, t.x, t.y, tw, th, blur, (t.x1, t.y1, t.x2, t.y2))
        orc.blend_processed(want, proc, t, m)
    feed = orc.quantize_u8(src) if src_u8 else src
    # (a) ONE launch with every tile in ascending order (the static-mode final blend)
    got = base.copy()
    km.run_blend(p, got, p.blend_worklist(ids, offs, 1 if src_u8 else 4, True, B), feed, pool)
    assert np.array_equal(got, want)
    # (b) the same launch shared out over 3 participants (dist.upscale_static)
    got = base.copy()
    for i in (2, 0, 1):                               # any order: the shares own disjoint blocks
        km.run_blend(p, got, p.blend_worklist(ids, offs, 1 if src_u8 else 4, True, B, part=(i, 3)), feed, pool)
    assert np.array_equal(got, want)
    # (c) wave by wave (the progressive driver blends each wave with its own launch)
    got = base.copy()
    pos = {t: i for i, t in enumerate(ids)}
    for wave in p.waves():
        km.run_blend(p, got, p.blend_worklist(wave, np.array([offs[pos[t]] for t in wave]), 1 if src_u8 else 4, True, B), feed, pool)
    assert np.array_equal(got, want)
