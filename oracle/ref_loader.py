"""Import the REAL reference pixel path from /root/reference under stub packages.

TEST INFRASTRUCTURE ONLY, and only usable in the build container: /root/reference does
not exist on the GPU box.  Used by ``oracle/gen_golden.py`` to produce the fixtures in
``tests/golden/`` and by ``tests/test_oracle_vs_reference.py`` (skipped when the
reference tree is absent).  Nothing is copied: the reference files are loaded from where
they lie with ``importlib`` (the same trick the reference's own tests use,
``tests/test_static_mode.py:11-128``), with ComfyUI (``comfy``, ``nodes``, ``server``)
replaced by minimal stand-ins.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

def _default_root() -> str:
    """/root/reference in the build container; elsewhere the archive oracle/make_ref.py packed (oracle/_ref, git-ignored),
    unpacked into a temporary directory."""
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import make_ref
    return make_ref.staged_root() or "/root/reference"


REF_ROOT = os.environ.get("USDU_REFERENCE_ROOT") or _default_root()
PKG = "refpkg"


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "upscale", "tile_ops.py"))


def _mod(name: str, **attrs) -> types.ModuleType:
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def _load(name: str, rel: str) -> types.ModuleType:
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF_ROOT, rel))
    m = importlib.util.module_from_spec(spec)
    sys.modules[name] = m
    spec.loader.exec_module(m)
    return m


class FakeNodes:
    """Stand-in for ComfyUI's top-level ``nodes`` module: identity VAE and a sampler
    that applies an injected callable ``fn(pixels, seed, denoise) -> pixels``."""

    def __init__(self):
        self.fn = None
        outer = self

        class VAEEncode:
            def encode(self, vae, pixels):
                return ({"samples": pixels},)

        class VAEDecode:
            def decode(self, vae, samples):
                return (samples["samples"],)

        def common_ksampler(model, seed, steps, cfg, sampler_name, scheduler, positive, negative,
                            latent, denoise=1.0):
            return ({"samples": outer.fn(latent["samples"], seed, denoise)},)

        self.module = _mod("nodes", VAEEncode=VAEEncode, VAEDecode=VAEDecode,
                           common_ksampler=common_ksampler)


_loaded = None


def load():
    """Returns (tile_ops module, single_gpu module, FakeNodes)."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")

    class _Interrupt(Exception):
        pass

    samplers = _mod("comfy.samplers", KSampler=types.SimpleNamespace(SAMPLERS=["euler"], SCHEDULERS=["normal"]))
    mm = _mod("comfy.model_management", processing_interrupted=lambda: False,
              throw_exception_if_processing_interrupted=lambda: None,
              InterruptProcessingException=_Interrupt)
    _mod("comfy", samplers=samplers, model_management=mm)
    fake_nodes = FakeNodes()

    for p in (PKG, PKG + ".utils", PKG + ".upscale", PKG + ".upscale.modes"):
        m = _mod(p)
        m.__path__ = []  # mark as package
    _mod(PKG + ".utils.logging", debug_log=lambda *a, **k: None, log=lambda *a, **k: None)
    _load(PKG + ".utils.image", "utils/image.py")
    _load(PKG + ".utils.usdu_utils", "utils/usdu_utils.py")
    _load(PKG + ".utils.crop_model_patch", "utils/crop_model_patch.py")
    _load(PKG + ".upscale.conditioning", "upscale/conditioning.py")
    tile_ops = _load(PKG + ".upscale.tile_ops", "upscale/tile_ops.py")
    single = _load(PKG + ".upscale.modes.single_gpu", "upscale/modes/single_gpu.py")
    _loaded = (tile_ops, single, fake_nodes)
    return _loaded


def make_reference_node():
    """An object with the reference's TileOpsMixin + SingleGpuModeMixin methods."""
    tile_ops, single, fake_nodes = load()

    class RefNode(single.SingleGpuModeMixin, tile_ops.TileOpsMixin):
        pass

    return RefNode(), fake_nodes
