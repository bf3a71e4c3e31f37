"""Seeded sweep over the node's parameter space on small canvases: whole job on the GPU ==
oracle.process_single, bit for bit.  Exercises partial edge tiles, canvases smaller than a
tile (up-sampling crops, > 7-tap and > 15-tap paths), padding 0 / large padding, blur 0 / large
blur (ramp > padding), non-uniform tiles, non-multiple-of-4 widths (scalar cast paths, TMA box
clipping) and multi-frame batches."""
import numpy as np
import pytest
import torch

import usdu_oracle as orc
from __graft_entry__ import load_package
from inputs import make_input

load_package()
from comfyui_distributed_b200 import engine, planner  # noqa: E402
from comfyui_distributed_b200.denoise import T0Denoiser  # noqa: E402

pytestmark = pytest.mark.gpu


def _cases():
    rng = np.random.default_rng(20260921)
    out = []
    for i in range(40):
        W = int(rng.integers(24, 700))
        H = int(rng.integers(24, 500))
        tw = int(rng.choice([64, 72, 96, 128, 200, 256, 512]))
        th = int(rng.choice([64, 80, 128, 256, 384]))
        pad = int(rng.choice([0, 8, 16, 32, 64, 128]))
        blur = int(rng.choice([0, 1, 4, 8, 16, 40, 97]))
        uniform = bool(rng.integers(0, 2))
        B = int(rng.choice([1, 1, 1, 2, 5]))
        kind = ["noise", "smooth", "checker"][int(rng.integers(0, 3))]
        out.append((i, kind, B, H, W, tw, th, pad, blur, uniform))
    out += [(100, "noise", 1, 64, 48, 512, 512, 32, 8, True),      # canvas << tile: 544 -> 48 needs 68 taps (generic kernels)
            (101, "noise", 1, 37, 1021, 64, 64, 8, 8, True),       # odd width, wide and flat
            (102, "checker", 1, 515, 33, 128, 128, 16, 255, True),  # narrow, blur far larger than the canvas
            (103, "smooth", 17, 96, 120, 64, 64, 16, 8, True)]      # WAN-style 4n+1 frame batch
    return out


@pytest.mark.parametrize("case", _cases(), ids=lambda c: f"{c[0]}-{c[1]}-b{c[2]}-{c[4]}x{c[3]}-t{c[5]}x{c[6]}-p{c[7]}-m{c[8]}-{'u' if c[9] else 'n'}")
def test_whole_job_matches_oracle(case):
    i, kind, B, H, W, tw, th, pad, blur, uniform = case
    img = make_input(kind, i, B, H, W)
    seed, den = 1000 + i, 0.25 + 0.05 * (i % 10)
    ref = orc.process_single(img, orc.make_t0_denoiser(seed, den), tw, th, pad, blur, uniform)
    st = {}
    out = engine.upscale_single(torch.from_numpy(img).cuda(), T0Denoiser(seed, den), tw, th, pad, blur, uniform, stats=st)
    assert np.array_equal(out.cpu().numpy(), ref), (case, planner.get_plan(W, H, tw, th, pad, blur, uniform).fast)
