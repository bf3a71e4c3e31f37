"""The real-ComfyUI code path of the node (denoise.ComfySampler: VAEEncode -> common_ksampler ->
VAEDecode per tile position, per-tile conditioning + mask crop, tile-local model patches, cancel
polling) driven by stand-in `nodes` / `comfy` modules, checked bit-exactly against the oracle's
process_single with the same per-tile arithmetic.  Mirrors how the reference's own tests stub
ComfyUI (tests/test_static_mode.py:11-128)."""
import sys
import types

import numpy as np
import pytest
import torch

import usdu_oracle as orc
from __graft_entry__ import load_package
from inputs import make_input, make_mask

load_package()
from comfyui_distributed_b200.nodes import UltimateSDUpscaleDistributed  # noqa: E402

pytestmark = pytest.mark.gpu
SEED, DENOISE = 77, 0.4


class Interrupted(Exception):
    pass


class DiffSynthCnetPatch:
    def __init__(self, model_patch, vae, image, strength, inpaint_image=None, mask=None):
        self.model_patch, self.vae, self.image, self.strength = model_patch, vae, image, strength
        self.inpaint_image, self.mask = inpaint_image, mask
        self.encoded_image, self.encoded_image_size = None, (image.shape[1], image.shape[2])


class Model:
    def __init__(self, patch):
        self.model_options = {"transformer_options": {"patches": {"double_block": [patch]}}}
        self.patch = patch

    def clone(self):
        m = Model(self.patch)
        return m


@pytest.fixture
def comfy(monkeypatch):
    """Stand-in ComfyUI: the 'sampler' is the T0 arithmetic gated by the tile's cropped mask."""
    log = types.SimpleNamespace(polls=0, patch_sizes=[], interrupt_at=None, mask_devices=set())
    nodes = types.ModuleType("nodes")

    class VAEEncode:
        def encode(self, vae, px):
            return ({"samples": px},)

    class VAEDecode:
        def decode(self, vae, s):
            return (s["samples"],)

    def common_ksampler(model, seed, steps, cfg, sampler_name, scheduler, pos, neg, latent, denoise=1.0):
        x = latent["samples"]
        g = torch.Generator().manual_seed(int(seed))
        noise = torch.rand(tuple(x.shape), generator=g, dtype=torch.float32).to(x.device)
        d = np.float32(denoise)
        y = torch.clamp(x * float(np.float32(1.0) - d) + noise * float(d), 0.0, 1.0)
        m = pos[0][1]["mask"]
        log.mask_devices.add(m.device.type)
        log.patch_sizes.append(tuple(model.patch.image.shape[1:3]))
        return ({"samples": torch.where(m.to(x.device)[..., None] >= 0.5, y, x)},)

    nodes.VAEEncode, nodes.VAEDecode, nodes.common_ksampler = VAEEncode, VAEDecode, common_ksampler
    comfy_pkg = types.ModuleType("comfy")
    mm = types.ModuleType("comfy.model_management")

    def poll():
        log.polls += 1
        if log.interrupt_at is not None and log.polls > log.interrupt_at:
            raise Interrupted()

    mm.throw_exception_if_processing_interrupted = poll
    comfy_pkg.model_management = mm
    monkeypatch.setitem(sys.modules, "nodes", nodes)
    monkeypatch.setitem(sys.modules, "comfy", comfy_pkg)
    monkeypatch.setitem(sys.modules, "comfy.model_management", mm)
    return log


def _oracle(img, mask, tw, th, pad, blur, uniform):
    H, W = img.shape[1:3]
    t0 = orc.make_t0_denoiser(SEED, DENOISE)
    q = [orc.quantize_u8(mask[b]) for b in range(mask.shape[0])]

    def fn(tile, t):
        y = t0(tile, t)
        m = np.stack([orc.crop_mask_u8(qb, (t.x1, t.y1, t.x1 + t.ew, t.y1 + t.eh), (W, H), (t.pw, t.ph)) for qb in q])
        m = m.astype(np.float32) / np.float32(255)
        return np.where(m[..., None] >= 0.5, y, tile)

    return orc.process_single(img, fn, tw, th, pad, blur, uniform)


@pytest.mark.parametrize("B,H,W,tile,pad,blur,uniform,where", [
    (1, 700, 900, 256, 32, 8, True, "cpu"), (1, 600, 520, 256, 16, 4, False, "cuda"), (5, 300, 420, 128, 16, 8, True, "cpu")])
def test_node_through_comfy_sampler_with_mask_and_model_patch(comfy, B, H, W, tile, pad, blur, uniform, where):
    img = make_input("smooth", 5, B, H, W)
    mask = make_mask("blob", 9, B, H // 4, W // 4)
    control = torch.rand(1, H // 2, W // 2, 3)
    patch = DiffSynthCnetPatch("mp", None, control, 1.0)
    cond = [[torch.zeros(1, 77, 8), {"mask": torch.from_numpy(mask), "pooled_output": torch.zeros(1, 8)}]]
    node = UltimateSDUpscaleDistributed()
    (out,) = node.run(torch.from_numpy(img).to(where), Model(patch), cond, cond, None, SEED, 20, 8.0, "euler", "normal",
                      DENOISE, tile, tile, pad, blur, uniform, False)
    want = _oracle(img, mask, tile, tile, pad, blur, uniform)
    assert out.device.type == where
    assert np.array_equal(out.cpu().numpy(), want)
    _, _, plan = orc.make_plan(W, H, tile, tile, pad, uniform)
    assert comfy.polls == len(plan)                                     # one cancel poll per tile position
    assert comfy.mask_devices == {"cuda"}
    from comfyui_distributed_b200.conditioning import scale_region
    sizes = set()
    for t in plan:
        x1, y1, x2, y2 = scale_region((t.x1, t.y1, t.x1 + t.ew, t.y1 + t.eh), (W, H), (W // 2, H // 2))
        sizes.add((y2 - y1, x2 - x1))
    assert set(comfy.patch_sizes) == sizes                              # the sampler saw tile-local control images
    assert patch.image is control or torch.equal(patch.image, control)   # ... and the patch is whole again


def test_user_cancel_propagates(comfy):
    comfy.interrupt_at = 3
    img = make_input("noise", 1, 1, 512, 512)
    cond = [[torch.zeros(1, 77, 8), {"mask": torch.ones(1, 64, 64)}]]
    patch = DiffSynthCnetPatch("mp", None, torch.rand(1, 64, 64, 3), 1.0)
    with pytest.raises(Interrupted):
        UltimateSDUpscaleDistributed().run(torch.from_numpy(img), Model(patch), cond, cond, None, SEED, 20, 8.0, "euler",
                                           "normal", DENOISE, 128, 128, 16, 4, True, False)
    assert comfy.polls == 4
