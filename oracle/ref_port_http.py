"""Cost-faithful port of the reference's N-participant static mode over HTTP + PNG --
BASELINE INFRASTRUCTURE ONLY (bench.py --impl reference --gpus N, N > 1).

The reference's "distributed" tile path is N ComfyUI processes talking aiohttp on
localhost.  /root/reference cannot travel to the GPU box and needs a
live ComfyUI, so this file restates that path with the same moving parts and the same
per-tile costs, on the host cores:

  master  (this process)   aiohttp server; pending-tile queue; pulls tiles itself; drains
                           the worker results; final blend of worker tiles in ascending
                           tile id                  upscale/modes/static.py:371-570
  workers (N-1 processes)  own u8 canvas + own masks; loop: POST request_image -> full-canvas
                           fp32 round trip -> crop/LANCZOS -> sampler -> local blend -> PNG
                           (compress_level 0) -> multipart POST submit_tiles, flushed every
                           MAX_BATCH tiles           static.py:191-314, worker_comms.py:16-188
  routes                   /distributed/request_image, /submit_tiles, /heartbeat,
                           /job_status               api/usdu_routes.py:16-228

Pixel work is done by oracle/ref_port.py (bit-identical to the reference).  A full 8K job
takes tens of minutes on CPU, so bench_sample() runs a BOUNDED sample -- the first
participants x tiles_per_participant tiles of the grid on the full-size canvas -- and
extrapolates per phase; the JSON line says so.
"""
from __future__ import annotations

import asyncio
import json
import multiprocessing as mp
import os
import sys
import threading
import time
from typing import Dict, List

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_port  # noqa: E402
import usdu_oracle as orc  # noqa: E402

MAX_BATCH = 20          # utils/constants.py:43 (COMFYUI_MAX_BATCH)


def _canvas(B, H, W):
    g = torch.Generator().manual_seed(0)
    return torch.floor(torch.rand(B, H, W, 3, generator=g) * 255) / 255


# --------------------------------------------------------------------------------------
# worker process
# --------------------------------------------------------------------------------------
def _worker(url: str, wid: str, cfg: dict, ret):
    import aiohttp
    torch.set_num_threads(max(1, (os.cpu_count() or 2) // cfg["participants"]))
    import urllib.request
    B, H, W, tile, pad, blur = cfg["workload"]
    port = ref_port.RefPort(W, H, tile, tile, pad, blur, True)
    image = _canvas(B, H, W)
    # a reference worker is an already-running ComfyUI process: report in, wait for the job to start
    urllib.request.urlopen(f"{url}/distributed/hello?worker_id={wid}").read()
    while json.loads(urllib.request.urlopen(f"{url}/distributed/go").read())["go"] is False:
        time.sleep(0.01)
    t_start = time.perf_counter()
    frames = [ref_port.to_pil(image[b:b + 1], 0).copy() for b in range(B)]          # static.py:209-212
    masks = {tid: port.feather(*port.grid[tid]) for tid in cfg["tile_ids"]}           # static.py:213-215 (sample: only the sampled tiles)
    denoise = ref_port.torch_t0(cfg["seed"], cfg["denoise"])
    t_setup = time.perf_counter() - t_start

    async def main():
        pending, done = [], 0
        t_png = t_http = 0.0
        async with aiohttp.ClientSession() as s:
            for _ in range(200):                                                      # _poll_job_ready, static.py:33-47
                async with s.get(f"{url}/distributed/job_status", params={"multi_job_id": "job"}) as r:
                    if (await r.json()).get("ready"):
                        break
                await asyncio.sleep(0.05)

            async def flush(final: bool):
                nonlocal pending, t_http
                t0 = time.perf_counter()
                form = aiohttp.FormData()
                form.add_field("multi_job_id", "job")
                form.add_field("worker_id", wid)
                form.add_field("is_last", "true" if final else "false")
                form.add_field("tiles_metadata", json.dumps([m for m, _ in pending]))
                for i, (_, png) in enumerate(pending):
                    form.add_field(f"tile_{i}", png, filename=f"tile_{i}.png", content_type="image/png")
                async with s.post(f"{url}/distributed/submit_tiles", data=form) as r:
                    await r.read()
                t_http += time.perf_counter() - t0
                pending = []

            while True:
                t0 = time.perf_counter()
                async with s.post(f"{url}/distributed/request_image", json={"multi_job_id": "job", "worker_id": wid}) as r:
                    tid = (await r.json()).get("tile_idx")
                t_http += time.perf_counter() - t0
                if tid is None:
                    break
                kept: Dict[int, tuple] = {}
                port.run_tiles(frames, [tid], masks, denoise, keep=kept)                 # static.py:242-280
                tile, x1, y1, ew, eh = kept[tid]
                t0 = time.perf_counter()
                for b in range(B):
                    png = ref_port.encode_tile_png(tile[b:b + 1])                        # worker_comms.py:30-33
                    pending.append(({"tile_idx": tid, "x": x1, "y": y1, "extracted_width": ew, "extracted_height": eh,
                                     "batch_idx": b, "global_idx": b * len(port.grid) + tid}, png))
                t_png += time.perf_counter() - t0
                done += 1
                t0 = time.perf_counter()
                async with s.post(f"{url}/distributed/heartbeat", json={"multi_job_id": "job", "worker_id": wid}) as r:
                    await r.read()
                t_http += time.perf_counter() - t0
                if len(pending) >= MAX_BATCH:
                    await flush(False)
            await flush(True)
        return done, t_png, t_http

    done, t_png, t_http = asyncio.run(main())
    ret.put({"worker": wid, "tiles": done, "setup_s": t_setup, "png_s": t_png, "http_s": t_http, **port.timer.t})


# --------------------------------------------------------------------------------------
# master
# --------------------------------------------------------------------------------------
def run_job(workload, seed, denoise, participants: int, tile_ids: List[int]) -> dict:
    from aiohttp import web
    B, H, W, tile, pad, blur = workload
    port_m = ref_port.RefPort(W, H, tile, tile, pad, blur, True)
    pending = list(tile_ids)
    lock = threading.Lock()
    results: Dict[int, tuple] = {}
    workers_done = set()
    state = {"ready": False, "go": False, "decode_s": 0.0}

    pulls: Dict[str, List[int]] = {}
    hello = set()

    async def request_image(req):
        body = await req.json()
        with lock:
            tid = pending.pop(0) if pending else None
            if tid is not None:
                pulls.setdefault(body["worker_id"], []).append(tid)
        return web.json_response({"tile_idx": tid, "estimated_remaining": len(pending), "batched_static": True})

    async def hello_route(req):
        hello.add(req.query["worker_id"])
        return web.json_response({"ok": True})

    async def go_route(req):
        return web.json_response({"go": state["go"]})

    async def submit_tiles(req):
        t0 = time.perf_counter()
        form = await req.post()
        meta = json.loads(form["tiles_metadata"])
        for i, m in enumerate(meta):
            img = ref_port.decode_tile_png(form[f"tile_{i}"].file.read())               # payload_parsers.py:7-64
            results[(m["tile_idx"], m["batch_idx"])] = (img, m)
        if form["is_last"] == "true":
            workers_done.add(form["worker_id"])
        state["decode_s"] += time.perf_counter() - t0
        return web.json_response({"status": "success"})

    async def heartbeat(req):
        return web.json_response({"status": "success"})

    async def job_status(req):
        return web.json_response({"ready": state["ready"]})

    app = web.Application(client_max_size=1 << 30)
    app.add_routes([web.post("/distributed/request_image", request_image), web.post("/distributed/submit_tiles", submit_tiles),
                    web.post("/distributed/heartbeat", heartbeat), web.get("/distributed/job_status", job_status),
                    web.get("/distributed/hello", hello_route), web.get("/distributed/go", go_route)])
    loop = asyncio.new_event_loop()
    runner = web.AppRunner(app)
    started = threading.Event()
    addr = {}

    def serve():
        asyncio.set_event_loop(loop)
        loop.run_until_complete(runner.setup())
        site = web.TCPSite(runner, "127.0.0.1", 0)
        loop.run_until_complete(site.start())
        addr["port"] = site._server.sockets[0].getsockname()[1]
        started.set()
        loop.run_forever()

    th = threading.Thread(target=serve, daemon=True)
    th.start()
    started.wait()
    url = f"http://127.0.0.1:{addr['port']}"
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    cfg = {"workload": workload, "seed": seed, "denoise": denoise, "participants": participants, "tile_ids": tile_ids}
    procs = [ctx.Process(target=_worker, args=(url, f"w{i}", cfg, ret)) for i in range(1, participants)]
    for p in procs:
        p.start()
    torch.set_num_threads(max(1, (os.cpu_count() or 2) // participants))
    image = _canvas(B, H, W)
    while len(hello) < participants - 1:            # all worker processes are up (not part of the job)
        time.sleep(0.01)
    state["go"] = True
    t_wall0 = time.perf_counter()
    t0 = time.perf_counter()
    frames = [ref_port.to_pil(image[b:b + 1], 0).copy() for b in range(B)]              # static.py:382-385
    t_q0 = time.perf_counter() - t0
    masks = {tid: port_m.feather(*port_m.grid[tid]) for tid in tile_ids}                  # static.py:398-400
    denoise_fn = ref_port.torch_t0(seed, denoise)
    state["ready"] = True
    mine = []
    t_loop0 = time.perf_counter()
    while True:                                                                          # master work loop, static.py:406-448
        with lock:
            tid = pending.pop(0) if pending else None
        if tid is None:
            break
        port_m.run_tiles(frames, [tid], masks, denoise_fn)
        mine.append(tid)
    t_own = time.perf_counter() - t_loop0
    t0 = time.perf_counter()
    while len(workers_done) < participants - 1:                                          # collection loop, static.py:316-369
        time.sleep(0.01)
    t_wait = time.perf_counter() - t0
    t0 = time.perf_counter()
    for (tid, b) in sorted(results):                                                     # static.py:521-553
        img, m = results[(tid, b)]
        frames[b] = port_m.blend(frames[b], img, m["x"], m["y"], m["extracted_width"], m["extracted_height"], masks[tid])
    t_final = time.perf_counter() - t0
    t0 = time.perf_counter()
    out = torch.cat([ref_port.to_tensor(f) for f in frames], dim=0)                      # static.py:556-564
    t_result = time.perf_counter() - t0
    wall = time.perf_counter() - t_wall0
    wstats = [ret.get(timeout=60) for _ in procs]
    for p in procs:
        p.join(timeout=30)
    loop.call_soon_threadsafe(loop.stop)
    return {"wall_s": wall, "q0_s": t_q0, "own_loop_s": t_own, "wait_workers_s": t_wait, "final_blend_s": t_final,
            "result_s": t_result, "decode_s": state["decode_s"], "master_tiles": mine, "worker_tiles": sorted({t for t, _ in results}),
            "mask_s_per_tile": port_m.timer.t.get("mask", 0.0) / max(len(tile_ids), 1), "master_phases": dict(port_m.timer.t),
            "pulls": pulls,
            "workers": wstats, "output": out}


def bench_sample(workload, seed, denoise, participants: int, tiles_per_participant: int = 1) -> dict:
    """Bounded sample + per-phase extrapolation to the full grid (see module docstring)."""
    B, H, W, tile, pad, blur = workload
    T = len(orc.calculate_tiles(W, H, orc.round_to_multiple(tile), orc.round_to_multiple(tile)))
    n = min(T, participants * tiles_per_participant)
    r = run_job(workload, seed, denoise, participants, list(range(n)))
    n_master, n_worker = max(len(r["master_tiles"]), 1), max(len(r["worker_tiles"]), 1)
    per_tile_parallel = r["own_loop_s"] / n_master                 # one participant's cost per tile it processes
    per_tile_final = r["final_blend_s"] / n_worker                 # master's serial blend per worker tile
    masks_all = r["mask_s_per_tile"] * T                           # every participant builds ALL T masks up front
    est = (r["q0_s"] + masks_all + per_tile_parallel * (T / participants) + per_tile_final * T * (participants - 1) / participants
           + r["result_s"])
    mp_total = B * H * W / 1e6
    return {"value": mp_total / est, "unit": "MP/s", "cores": os.cpu_count(), "kind": "port",
            "sample": f"HTTP+PNG static mode, master + {participants - 1} local worker processes, first {n} of {T} tiles on the "
                      f"full canvas (oracle/ref_port_http.py); extrapolated {est:.0f}s/job = q0 {r['q0_s']:.1f} + masks {masks_all:.1f} + "
                      f"{T / participants:.1f} tiles x {per_tile_parallel:.2f}s + {T * (participants - 1) / participants:.0f} worker tiles x "
                      f"{per_tile_final:.2f}s final blend + result {r['result_s']:.1f}",
            "sample_wall_s": round(r["wall_s"], 2), "participants": participants,
            "phases_s": {k: round(v, 3) for k, v in r.items() if k.endswith("_s")},
            "worker_phases_s": [{k: (round(v, 3) if isinstance(v, float) else v) for k, v in w.items()} for w in r["workers"]]}


if __name__ == "__main__":
    print(json.dumps(bench_sample((1, 1100, 1300, 512, 32, 8), 123, 0.5, participants=3, tiles_per_participant=3)))
