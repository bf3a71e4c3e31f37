"""Run the REAL reference's static (multi-worker) mode in this container: master + workers over
a real aiohttp server on 127.0.0.1 with the reference's own PNG-multipart transport.

TEST INFRASTRUCTURE ONLY, build container only (/root/reference does not exist on the GPU box).
Nothing is copied: the reference's files are loaded from where they lie (same stub-package trick as
oracle/ref_loader.py and the reference's tests/test_static_mode.py:11-128); ComfyUI (`server`,
`execution`, `comfy`, `nodes`) is replaced by minimal stand-ins, the sampler by the T0 arithmetic.

`run_static(...)` returns the master's result AND the tile assignment that actually happened (the
reference's workers PULL tile ids over HTTP, so it differs from run to run); `oracle/gen_golden.py`
stores both, and tests replay the recorded assignment through `usdu_oracle.replay_static` and through
the CUDA path.  Files exercised: nodes/distributed_upscale.py (run / process_master / process_worker),
upscale/modes/static.py, upscale/worker_comms.py, upscale/result_collector.py, upscale/job_store.py,
api/usdu_routes.py, upscale/payload_parsers.py, upscale/tile_ops.py, utils/image.py.
"""
from __future__ import annotations

import asyncio
import importlib.util
import json
import os
import socket
import sys
import threading
import time
import types
from typing import Dict, List

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import usdu_oracle as orc  # noqa: E402

def _default_root() -> str:
    """/root/reference in the build container; elsewhere the archive oracle/make_ref.py packed (oracle/_ref, git-ignored),
    unpacked into a temporary directory."""
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import make_ref
    return make_ref.staged_root() or "/root/reference"


REF_ROOT = os.environ.get("USDU_REFERENCE_ROOT") or _default_root()
PKG = "refstatic"

ORDER = [
    ("utils.constants", "utils/constants.py"), ("utils.config", "utils/config.py"), ("utils.network", "utils/network.py"),
    ("utils.image", "utils/image.py"), ("utils.usdu_utils", "utils/usdu_utils.py"),
    ("utils.crop_model_patch", "utils/crop_model_patch.py"), ("utils.async_helpers", "utils/async_helpers.py"),
    ("upscale.conditioning", "upscale/conditioning.py"), ("upscale.job_models", "upscale/job_models.py"),
    ("upscale.job_store", "upscale/job_store.py"), ("upscale.job_timeout", "upscale/job_timeout.py"),
    ("upscale.payload_parsers", "upscale/payload_parsers.py"), ("utils.usdu_managment", "utils/usdu_managment.py"),
    ("upscale.job_state", "upscale/job_state.py"), ("upscale.tile_ops", "upscale/tile_ops.py"),
    ("upscale.result_collector", "upscale/result_collector.py"), ("upscale.worker_comms", "upscale/worker_comms.py"),
    ("upscale.modes.single_gpu", "upscale/modes/single_gpu.py"), ("upscale.modes.static", "upscale/modes/static.py"),
    ("upscale.modes.dynamic", "upscale/modes/dynamic.py"), ("api.usdu_routes", "api/usdu_routes.py"),
    ("nodes.distributed_upscale", "nodes/distributed_upscale.py"),
]


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "upscale", "modes", "static.py"))


def _mod(name: str, **attrs) -> types.ModuleType:
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def _free_port() -> int:
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class _Env:
    """One event loop thread + one aiohttp site hosting the reference's USDU routes."""

    def __init__(self):
        from aiohttp import web
        self.loop = asyncio.new_event_loop()
        self.thread = threading.Thread(target=self._run_loop, daemon=True)
        self.thread.start()
        self.port = _free_port()
        routes = web.RouteTableDef()
        inst = types.SimpleNamespace(routes=routes, loop=self.loop, port=self.port, address="127.0.0.1",
                                     client_id=None, prompt_queue=None, last_node_id=None,
                                     send_sync=lambda *a, **k: None)
        self.saved = {k: sys.modules.get(k) for k in ("server", "execution", "comfy", "comfy.samplers",
                                                     "comfy.model_management", "comfy.utils", "nodes")}
        _mod("server", PromptServer=types.SimpleNamespace(instance=inst))
        _mod("execution", validate_prompt=None, PromptQueue=object)

        class _Interrupt(Exception):
            pass

        samplers = _mod("comfy.samplers", KSampler=types.SimpleNamespace(SAMPLERS=["euler"], SCHEDULERS=["normal"]))
        mm = _mod("comfy.model_management", processing_interrupted=lambda: False,
                  throw_exception_if_processing_interrupted=lambda: None, InterruptProcessingException=_Interrupt)
        cu = _mod("comfy.utils", ProgressBar=lambda *a, **k: types.SimpleNamespace(update=lambda *a, **k: None,
                                                                                update_absolute=lambda *a, **k: None))
        _mod("comfy", samplers=samplers, model_management=mm, utils=cu)
        self.sampler = None                      # fn(pixels [B,h,w,3] torch, seed, denoise) -> pixels
        outer = self

        class VAEEncode:
            def encode(self, vae, pixels):
                return ({"samples": pixels},)

        class VAEDecode:
            def decode(self, vae, samples):
                return (samples["samples"],)

        def common_ksampler(model, seed, steps, cfg, sampler_name, scheduler, positive, negative, latent, denoise=1.0):
            return ({"samples": outer.sampler(latent["samples"], seed, denoise)},)

        _mod("nodes", VAEEncode=VAEEncode, VAEDecode=VAEDecode, common_ksampler=common_ksampler)
        for p in (PKG, PKG + ".utils", PKG + ".upscale", PKG + ".upscale.modes", PKG + ".api", PKG + ".nodes"):
            _mod(p).__path__ = []
        _mod(PKG + ".utils.logging", debug_log=lambda *a, **k: None, log=lambda *a, **k: None)
        self.mods = {}
        for name, rel in ORDER:
            spec = importlib.util.spec_from_file_location(f"{PKG}.{name}", os.path.join(REF_ROOT, rel))
            m = importlib.util.module_from_spec(spec)
            sys.modules[f"{PKG}.{name}"] = m
            spec.loader.exec_module(m)
            self.mods[name] = m
        # never touch the read-only reference tree: the config file lives nowhere
        self.mods["utils.config"].CONFIG_FILE = os.path.join("/nonexistent", "gpu_config.json")
        app = web.Application(client_max_size=1 << 30)
        app.add_routes(routes)
        self.runner = web.AppRunner(app)
        self._call(self.runner.setup())
        self._call(web.TCPSite(self.runner, "127.0.0.1", self.port).start())
        self.node_cls = self.mods["nodes.distributed_upscale"].UltimateSDUpscaleDistributed

    def _run_loop(self):
        asyncio.set_event_loop(self.loop)
        self.loop.run_forever()

    def _call(self, coro, timeout=60):
        return asyncio.run_coroutine_threadsafe(coro, self.loop).result(timeout)

    def close(self):
        try:
            net = self.mods.get("utils.network")
            if net is not None and hasattr(net, "cleanup_client_session"):
                self._call(net.cleanup_client_session())
            self._call(self.runner.cleanup())
        finally:
            self.loop.call_soon_threadsafe(self.loop.stop)
            self.thread.join(timeout=10)
            for k, v in self.saved.items():
                if v is None:
                    sys.modules.pop(k, None)
                else:
                    sys.modules[k] = v
            for k in [k for k in sys.modules if k == PKG or k.startswith(PKG + ".")]:
                del sys.modules[k]


def torch_t0(pixels: torch.Tensor, seed: int, denoise: float) -> torch.Tensor:
    """The T0 arithmetic (oracle.make_t0_denoiser) on torch CPU tensors."""
    g = torch.Generator().manual_seed(int(seed))
    noise = torch.rand(tuple(pixels.shape), generator=g, dtype=torch.float32)
    d = np.float32(denoise)
    return torch.clamp(pixels.float() * float(np.float32(1.0) - d) + noise * float(d), 0.0, 1.0)


def run_static(image: np.ndarray, n_workers: int, tile: int, padding: int, mask_blur: int, uniform: bool,
               seed: int, denoise: float, job_id: str = "job1", timeout: float = 600.0, master_delay: float = 0.0,
               max_tiles: int = 0):
    """-> (result fp32 [B,H,W,3] of the reference's master, assignment: list over participants
    (master first, then w1..wN) of the tile ids each one processed, in processing order).
    master_delay: seconds the master sleeps before each of its tiles, so that the workers (which
    start later and talk HTTP) pull a fair share of the queue.  max_tiles: every participant sees only the first
    max_tiles tiles of the grid (the reference code is untouched: its calculate_tiles is wrapped on the node OBJECT)."""
    env = _Env()
    try:
        env.sampler = torch_t0
        B, H, W, _ = image.shape
        _, _, plan = orc.make_plan(W, H, tile, tile, padding, uniform)
        by_origin = {(t.x, t.y): t.idx for t in plan}        # crop windows can coincide (partial last row), origins cannot
        names = ["master"] + [f"w{i + 1}" for i in range(n_workers)]
        log: Dict[str, List[int]] = {n: [] for n in names}

        def make_node(name):
            node = env.node_cls()
            inner = node.extract_batch_tile_with_padding       # called once per processed tile position (static.py:74)

            def spy(upscaled_image, tx, ty, *rest):
                if name == "master" and master_delay > 0:
                    time.sleep(master_delay)
                log[name].append(by_origin[(int(tx), int(ty))])
                return inner(upscaled_image, tx, ty, *rest)

            node.extract_batch_tile_with_padding = spy
            if max_tiles > 0:                                   # bench.py's bounded sample: the job is the first max_tiles tiles
                full = node.calculate_tiles
                node.calculate_tiles = lambda *a, **k: full(*a, **k)[:max_tiles]
            return node

        x = torch.from_numpy(image)
        cond = [[torch.zeros(1, 77, 8), {}]]
        common = (None, cond, cond, None, seed, 20, 8.0, "euler", "normal", denoise, tile, tile, padding, mask_blur,
                  uniform, False)
        workers_json = json.dumps(names[1:])
        results, errors = {}, {}

        def participant(name):
            try:
                node = make_node(name)
                if name == "master":
                    results[name] = node.run(x.clone(), *common, multi_job_id=job_id, is_worker=False,
                                             enabled_worker_ids=workers_json)
                else:
                    results[name] = node.run(x.clone(), *common, multi_job_id=job_id, is_worker=True,
                                             master_url=f"http://127.0.0.1:{env.port}", worker_id=name,
                                             enabled_worker_ids=workers_json)
            except BaseException as e:      # noqa: BLE001 -- reported to the caller below
                errors[name] = e

        threads = [threading.Thread(target=participant, args=(n,), daemon=True) for n in names]
        threads[0].start()                  # the master creates the job; workers poll until it exists
        for t in threads[1:]:
            t.start()
        for t in threads:
            t.join(timeout)
        if errors:
            raise RuntimeError(f"reference static run failed: {errors}")
        if any(t.is_alive() for t in threads):
            raise TimeoutError("reference static run did not finish")
        out = results["master"][0].numpy()
        for n in names[1:]:                 # workers hand their input back unchanged (static.py:314)
            assert np.array_equal(results[n][0].numpy(), image), n
        return out, [log[n] for n in names]
    finally:
        env.close()


if __name__ == "__main__":
    sys.path.insert(0, os.path.join(os.path.dirname(HERE), "tests"))
    from inputs import make_input
    img = make_input("noise", 3, 1, 520, 700)
    res, asg = run_static(img, 1, 256, 32, 8, True, 9, 0.5)
    print("assignment", asg)
    ref = orc.replay_static(img, orc.make_t0_denoiser(9, 0.5), 256, 256, 32, 8, True, asg)
    print("replay_static == real reference:", np.array_equal(ref, res), float(np.abs(ref - res).max()))
