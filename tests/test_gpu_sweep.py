"""Seeded sweep over the node's parameter space on small canvases: whole job on the GPU ==
oracle.process_single == the REAL reference's process_single_gpu (digests made by oracle/gen_golden.py from
/root/reference: tests/golden/sweep_ref_digests.json), bit for bit.  Exercises partial edge tiles, canvases smaller than a
tile (up-sampling crops, > 7-tap and > 15-tap paths), padding 0 / large padding, blur 0 / large
blur (ramp > padding), non-uniform tiles, non-multiple-of-4 widths (scalar cast paths, TMA box
clipping) and multi-frame batches."""
import hashlib
import json
import os

import numpy as np
import pytest
import torch

import usdu_oracle as orc
from __graft_entry__ import load_package
from inputs import make_input, sweep_cases, sweep_sampler

load_package()
from comfyui_distributed_b200 import engine, planner  # noqa: E402
from comfyui_distributed_b200.denoise import T0Denoiser  # noqa: E402

pytestmark = pytest.mark.gpu
DIGESTS = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "sweep_ref_digests.json")))["digests"]


@pytest.fixture(params=["tensor-core", "integer-pipe"])
def kernels(request):
    """Default kernel choice (tensor-core where the plan allows, else the integer-pipe fast ones, else generic) and the
    same sweep with the tensor-core kernels switched off."""
    engine.FORCE_NO_MMA = request.param == "integer-pipe"
    yield request.param
    engine.FORCE_NO_MMA = False


@pytest.mark.parametrize("case", sweep_cases(), ids=lambda c: f"{c[0]}-{c[1]}-b{c[2]}-{c[4]}x{c[3]}-t{c[5]}x{c[6]}-p{c[7]}-m{c[8]}-{'u' if c[9] else 'n'}")
def test_whole_job_matches_oracle_and_reference(case, kernels):
    i, kind, B, H, W, tw, th, pad, blur, uniform = case
    img = make_input(kind, i, B, H, W)
    seed, den = sweep_sampler(i)
    ref = orc.process_single(img, orc.make_t0_denoiser(seed, den), tw, th, pad, blur, uniform)
    st = {}
    out = engine.upscale_single(torch.from_numpy(img).cuda(), T0Denoiser(seed, den), tw, th, pad, blur, uniform, stats=st)
    assert np.array_equal(out.cpu().numpy(), ref), (case, planner.get_plan(W, H, tw, th, pad, blur, uniform).fast)
    q = np.round(out.cpu().numpy() * 255).astype(np.uint8)
    assert hashlib.sha256(q.tobytes()).hexdigest() == DIGESTS[str(i)]          # what the reference itself produced
