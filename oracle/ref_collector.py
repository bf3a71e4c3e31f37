"""The REAL reference's DistributedCollector arithmetic, loaded from /root/reference under stub packages
(TEST INFRASTRUCTURE ONLY, build container only): worker side `tensor_to_pil` + PNG level 0 + base64
(nodes/collector.py:84-103), master side `_decode_canonical_png_tensor` (api/job_routes.py:104-132) and
`_reorder_and_combine_tensors` (nodes/collector.py:193-236).  Used to pin `usdu_oracle.collector_combine`
(tests/test_collector_vs_reference.py) and to make tests/golden/collector_ref.json."""
from __future__ import annotations

import base64
import importlib.util
import io
import os
import sys
import types
from typing import Dict, Sequence

import torch

REF_ROOT = os.environ.get("USDU_REFERENCE_ROOT", "/root/reference")
PKG = "refcollector"


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "nodes", "collector.py"))


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def _load(name, rel):
    spec = importlib.util.spec_from_file_location(f"{PKG}.{name}", os.path.join(REF_ROOT, rel))
    m = importlib.util.module_from_spec(spec)
    sys.modules[f"{PKG}.{name}"] = m
    spec.loader.exec_module(m)
    return m


_loaded = None


def load():
    """-> (collector module, job_routes module, image module)"""
    global _loaded
    if _loaded is not None:
        return _loaded
    from aiohttp import web
    saved = {k: sys.modules.get(k) for k in ("server", "execution", "comfy", "comfy.model_management", "comfy.utils")}
    inst = types.SimpleNamespace(routes=web.RouteTableDef(), loop=None, port=0, address="127.0.0.1")
    _mod("server", PromptServer=types.SimpleNamespace(instance=inst))
    _mod("execution")
    mm = _mod("comfy.model_management", processing_interrupted=lambda: False,
              throw_exception_if_processing_interrupted=lambda: None)
    cu = _mod("comfy.utils", ProgressBar=lambda *a, **k: types.SimpleNamespace(update=lambda *a, **k: None))
    _mod("comfy", model_management=mm, utils=cu)
    for p in (PKG, PKG + ".utils", PKG + ".api", PKG + ".nodes"):
        _mod(p).__path__ = []
    _mod(PKG + ".utils.logging", debug_log=lambda *a, **k: None, log=lambda *a, **k: None)
    # control-plane modules job_routes imports at module level but the decode path never calls
    _mod(PKG + ".api.queue_orchestration", ensure_distributed_state=lambda: None, orchestrate_distributed_execution=None)
    _mod(PKG + ".api.queue_request", parse_queue_request_payload=None)
    for name, rel in (("utils.constants", "utils/constants.py"), ("utils.config", "utils/config.py"),
                      ("utils.network", "utils/network.py"), ("utils.image", "utils/image.py"),
                      ("utils.audio_payload", "utils/audio_payload.py"), ("utils.async_helpers", "utils/async_helpers.py")):
        _load(name, rel)
    routes = _load("api.job_routes", "api/job_routes.py")
    collector = _load("nodes.collector", "nodes/collector.py")
    image = sys.modules[f"{PKG}.utils.image"]
    for k, v in saved.items():
        if v is None:
            sys.modules.pop(k, None)
        else:
            sys.modules[k] = v
    _loaded = (collector, routes, image)
    return _loaded


def worker_roundtrip(images: torch.Tensor) -> Dict[int, torch.Tensor]:
    """What the master holds for one worker after the reference's transport: {image index: [1,H,W,3]}."""
    collector, routes, image = load()
    out = {}
    for i in range(images.shape[0]):
        pil = image.tensor_to_pil(images[i:i + 1], 0)                    # collector.py:95
        buf = io.BytesIO()
        pil.save(buf, format="PNG", compress_level=0)                    # :97
        payload = "data:image/png;base64," + base64.b64encode(buf.getvalue()).decode("utf-8")   # :98-103
        out[i] = routes._decode_canonical_png_tensor(payload)            # job_routes.py:104-132
    return out


def combine(master_images: torch.Tensor, worker_images: Dict[str, torch.Tensor], worker_order: Sequence[str],
            delegate_only: bool = False) -> torch.Tensor:
    collector, _, _ = load()
    node = collector.DistributedCollectorNode()
    held = {str(w): worker_roundtrip(t) for w, t in worker_images.items()}
    return node._reorder_and_combine_tensors(held, list(worker_order), int(master_images.shape[0]),
                                             master_images.cpu(), bool(delegate_only), master_images)
