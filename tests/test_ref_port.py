"""The cost-faithful CPU port (oracle/ref_port.py, what bench.py times as the CPU baseline)
produces exactly the real reference's output (fixtures from oracle/gen_golden.py)."""
import json
import os

import numpy as np
import pytest
import torch

import ref_port
import usdu_oracle as orc
from inputs import make_input

G = os.path.join(os.path.dirname(__file__), "golden")
SINGLE = json.load(open(os.path.join(G, "single_index.json")))["cases"]


@pytest.mark.parametrize("case", [c for c in SINGLE if c["name"] in ("odd_b2", "nonuniform", "upsample", "b5_video")],
                         ids=lambda c: c["name"])
def test_port_matches_reference_golden(case):
    img = make_input(case["kind"], case["seed"], case["B"], case["H"], case["W"])
    t = {}
    res = ref_port.process_single(torch.from_numpy(img), ref_port.torch_t0(case["denoise_seed"], case["denoise"]),
                                  case["tile_w"], case["tile_h"], case["padding"], case["mask_blur"], case["uniform"],
                                  timer_out=t)
    ref = np.load(os.path.join(G, f"single_{case['name']}.npz"))["out"]
    assert np.array_equal(res.numpy(), orc.dequantize_u8(ref))
    assert t["tiles_done"] == t["tiles_total"] and t["blend"] > 0


def test_png_codec_round_trip():
    x = torch.rand(1, 40, 56, 3)
    img = ref_port.decode_tile_png(ref_port.encode_tile_png(x))
    assert np.array_equal(np.array(img), orc.quantize_u8(x[0].numpy()))
