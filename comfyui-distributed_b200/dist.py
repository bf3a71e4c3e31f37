"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL over NVLink 5 /
NVSwitch; gloo in the CPU tests).  Replaces the reference's HTTP + PNG transport:

* upscale/worker_comms.py:16-108  (PNG multipart POST of processed tiles)  and
  upscale/result_collector.py:36-182 (master drain loop)   -> u8 tiles stay in their owner's HBM (symmetric
                                                              memory); the blend kernels of all ranks read them
                                                              over NVLink and composite shares of the master's
                                                              canvas in place (fallback: all_gather of u8 tiles)
* upscale/worker_comms.py:124-188 (HTTP pull of tile ids)  -> static plan (planner.partition)
* nodes/collector.py:84-119 + api/job_routes.py:273-343 (base64 PNG per image)
                                                           -> all_gather of u8 images

Semantics kept (`replay_static`): every participant starts from the
quantised input, crops from ITS OWN progressive canvas, and the result is the master's
canvas with every worker tile blended on top in ascending tile id
(upscale/modes/static.py:521-553).  Workers return their input unchanged (:314).
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as td


def dist_info(group=None) -> Tuple[int, int]:
    if td.is_available() and td.is_initialized():
        return td.get_rank(group), td.get_world_size(group)
    return 0, 1


# --------------------------------------------------------------------------------------
# transport (device agnostic: NCCL on CUDA tensors, gloo on CPU tensors in the tests)
# --------------------------------------------------------------------------------------
def all_gather_bytes(payload: torch.Tensor, group=None, sizes: Optional[Sequence[int]] = None) -> Tuple[torch.Tensor, List[int]]:
    """All-gather variable-length u8 payloads.  Returns (buffer [world, cap], sizes).
    One padded all_gather_into_tensor, preceded by a size exchange (int64 all_gather) only
    when the caller cannot supply `sizes` (the tile paths know them from the plan)."""
    assert payload.dtype == torch.uint8 and payload.dim() == 1
    rank, world = dist_info(group)
    if world == 1:
        return payload.view(1, -1), [payload.numel()]
    if sizes is None:
        n = torch.tensor([payload.numel()], dtype=torch.int64, device=payload.device)
        got = torch.empty(world, dtype=torch.int64, device=payload.device)
        td.all_gather_into_tensor(got, n, group=group)
        sizes = got.tolist()
    sizes = [int(s) for s in sizes]
    cap = max(max(sizes), 1)
    cap = (cap + 15) // 16 * 16
    send = payload
    if payload.numel() != cap:
        send = torch.zeros(cap, dtype=torch.uint8, device=payload.device)
        send[: payload.numel()] = payload
    out = torch.empty((world, cap), dtype=torch.uint8, device=payload.device)
    td.all_gather_into_tensor(out.view(-1), send, group=group)
    return out, sizes


def tile_payload_layout(plan, assignment: Sequence[Sequence[int]], B: int):
    """Byte offset of every tile's u8 [B,ph,pw,3] block inside its owner's payload, in
    the owner's processing order.  -> ({tile: (rank, offset)}, payload bytes per rank)"""
    where: Dict[int, Tuple[int, int]] = {}
    sizes = []
    for r, tiles in enumerate(assignment):
        cur = 0
        for tid in tiles:
            t = plan.tiles[tid]
            where[tid] = (r, cur)
            cur += B * t.ph * t.pw * 3
            cur = (cur + 15) // 16 * 16
        sizes.append(cur)
    return where, sizes


# --------------------------------------------------------------------------------------
# peer transport: the master's blend kernel pulls worker tiles straight out of the workers'
# HBM over NVLink (no gather step, no staging copy; the transfer overlaps the blend math CTA
# by CTA).  The payload buffers are symmetric allocations (same size on every rank) whose
# device addresses are exchanged once by torch's symmetric-memory rendezvous; inside one
# unified virtual address space a tile of rank r is simply  own_base + (ptr[r] - ptr[own]) +
# offset, which the blend kernel's 64-bit source offsets already express.
# --------------------------------------------------------------------------------------
USE_PEER_BLEND = os.environ.get("USDU_PEER_BLEND", "1") != "0"
# ... and, with it, the final blend itself is shared out: the master's canvas is a symmetric
# allocation too, every rank composites its share of the canvas BLOCKS (all worker tiles, ascending
# id, same order inside every block) straight into the master's HBM, reading the tiles from their
# owners' HBM.  Blocks are owned by exactly one CTA of exactly one rank, so the result is the one
# the single launch on the master produces.
USE_SHARED_FINAL_BLEND = os.environ.get("USDU_SHARED_FINAL_BLEND", "1") != "0"


def peer_offsets(order: Sequence[int], where: Dict[int, Tuple[int, int]], ptrs: Sequence[int], own_rank: int) -> np.ndarray:
    """Byte offset, relative to this rank's payload base, of every tile of `order` inside its
    owner's payload buffer (ptrs[r] = device address of rank r's buffer as mapped HERE)."""
    base = int(ptrs[own_rank])
    return np.array([int(ptrs[where[t][0]]) - base + where[t][1] for t in order], dtype=np.int64)


class PeerPayload:
    """Symmetric u8 buffer + rendezvous handle, cached per (tag, bytes, device, group)."""

    _cache: Dict[tuple, Optional["PeerPayload"]] = {}

    def __init__(self, nbytes: int, device, group):
        import torch.distributed._symmetric_memory as symm
        self.buf = symm.empty(nbytes, dtype=torch.uint8, device=device)
        self.hdl = symm.rendezvous(self.buf, group if group is not None else td.group.WORLD)
        self.ptrs = [int(p) for p in self.hdl.buffer_ptrs]
        self.buf.zero_()

    def barrier(self, channel: int = 0):
        """Stream-ordered barrier over all ranks (signal pads in peer memory, system-scope
        release/acquire): kernels enqueued before it on any rank are complete and visible to
        kernels enqueued after it on every rank."""
        self.hdl.barrier(channel=channel, timeout_ms=20000)

    @classmethod
    def get(cls, nbytes: int, device, group, tag: str = "payload") -> Optional["PeerPayload"]:
        """Collective.  Returns None on EVERY rank if any rank cannot set the buffer up (no
        NVLink/P2P, symmetric memory unsupported) -- the caller then uses the NCCL all-gather."""
        if not (USE_PEER_BLEND and td.is_initialized() and td.get_backend(group) == "nccl"):
            return None
        key = (tag, int(nbytes), str(device), id(group))
        if key in cls._cache:
            return cls._cache[key]
        obj, ok = None, 1
        try:
            obj = PeerPayload(int(nbytes), device, group)
        except Exception as e:     # noqa: BLE001 -- any failure means "transport not available here"
            ok = 0
            cls.last_error = repr(e)
        flag = torch.tensor([ok], dtype=torch.int32, device=device)
        td.all_reduce(flag, op=td.ReduceOp.MIN, group=group)
        if int(flag.item()) == 0:
            obj = None
        if len(cls._cache) > 8:
            cls._cache.clear()
        cls._cache[key] = obj
        return obj

    last_error: Optional[str] = None


def final_blend_order(assignment: Sequence[Sequence[int]]) -> List[int]:
    """Tile ids of all NON-master participants in the order the master composites them
    (ascending tile id, upscale/modes/static.py:521-526)."""
    return sorted(t for r, tiles in enumerate(assignment) if r != 0 for t in tiles)


# --------------------------------------------------------------------------------------
# static mode, SPMD
# --------------------------------------------------------------------------------------
def upscale_static(image: torch.Tensor, denoiser, tile_width: int, tile_height: int, padding: int,
                   mask_blur: int, force_uniform_tiles: bool = True, group=None,
                   assignment: Optional[Sequence[Sequence[int]]] = None, all_ranks_result: bool = False,
                   stats: Optional[dict] = None) -> torch.Tensor:
    """Every rank calls this with the same (replicated) CUDA image, like the reference's
    workers which each re-execute the upstream graph.  Rank 0 returns the
    blended canvas; other ranks return `image` unchanged unless all_ranks_result."""
    from . import _native as nat
    from .engine import Canvas, DevicePlan, _require_cuda, _stream_ptr, run_progressive
    from .planner import get_plan

    _require_cuda(image, "image")
    rank, world = dist_info(group)
    B, H, W, _ = image.shape
    plan = get_plan(W, H, tile_width, tile_height, padding, mask_blur, force_uniform_tiles)
    asg = [list(a) for a in (assignment if assignment is not None else plan.partition(world))]
    if len(asg) != world:
        raise ValueError(f"assignment has {len(asg)} participants, world size is {world}")
    with torch.cuda.device(image.device):
        dp = DevicePlan.get(plan, image.device)
        from . import engine as _eng
        # payload layout follows every rank's processing order (wave by wave), so a wave's
        # packed u8 tiles land in one contiguous span
        proc = [_eng.processing_order(plan, a) for a in asg]
        where, sizes = tile_payload_layout(plan, proc, B)
        sizes = [max(sz, 16) for sz in sizes]
        graphed = (bool(getattr(denoiser, "cuda_graph_safe", False)) and _eng.USE_CUDA_GRAPHS and not all_ranks_result
                   and len(asg[rank]) > 0)          # a participant without tiles has nothing to capture
        payload = None
        peer = PeerPayload.get((max(sizes) + 255) // 256 * 256, image.device, group) if world > 1 else None
        shared = None     # symmetric canvases: every rank can composite into the master's
        if peer is not None and USE_SHARED_FINAL_BLEND and not all_ranks_result:
            shared = PeerPayload.get(B * H * Canvas.pitch_of(W), image.device, group, tag="canvas")
        cbuf = shared.buf.view(B, H, Canvas.pitch_of(W)) if shared is not None else None
        if world > 1:
            payload = (peer.buf[: sizes[rank]] if peer is not None
                       else _payload_buffer(sizes[rank], image.device))
        if graphed:      # this rank's wave loop (crop -> sampler -> local blend -> u8 pack) as one CUDA graph
            gw = _eng.GraphedWaves.get(dp, B, denoiser, _eng.PROFILE, order=asg[rank], payload=payload, where=where,
                                       canvas_buf=cbuf)
            canvas, base = gw.replay(image), None
        else:
            canvas = Canvas(dp, B, cbuf).load(image)
            base = canvas.clone() if (all_ranks_result and rank != 0) else None
            run_progressive(canvas, asg[rank], denoiser, payload=payload, where=where)
        if world > 1:
            produce_here = rank == 0 or all_ranks_result
            target, order = canvas, final_blend_order(asg)
            if produce_here and rank != 0:
                # rebuild the master's canvas: base + master tiles in the master's order
                target, order = base, list(asg[0]) + order
            if shared is not None:
                peer.barrier(0)                          # every payload is complete, the master's own tiles are blended
                canvas.blend(order, peer.buf, peer_offsets(order, where, peer.ptrs, rank), part=(rank, world),
                             canvas_ptr=shared.ptrs[0])
                peer.barrier(1)                          # every share has landed in the master's canvas
            elif peer is not None:
                peer.barrier(0)                          # every payload is complete and visible
                if produce_here:
                    target.blend(order, peer.buf, peer_offsets(order, where, peer.ptrs, rank))
                peer.barrier(1)                          # nobody refills its payload while it is being read
            else:
                gathered, _ = all_gather_bytes(payload, group, sizes=sizes)
                cap = gathered.shape[1]
                if produce_here:
                    offs = np.array([where[t][0] * cap + where[t][1] for t in order], dtype=np.int64)
                    target.blend(order, gathered.view(-1), offs)
            if produce_here:
                canvas = target
        produce = rank == 0 or all_ranks_result
        res = canvas.result() if produce else image
    if stats is not None:
        stats["gpu_launches"] = stats.get("gpu_launches", 0) + canvas.launches
        stats["algo_bytes"] = stats.get("algo_bytes", 0) + canvas.algo_bytes
        stats["tiles"] = len(plan.tiles)
        stats["tiles_this_rank"] = len(asg[rank])
        stats["conflict_free"] = plan.conflict_free(asg)
        stats["transport"] = "single" if world == 1 else ("nvlink peer loads" if peer is not None else "nccl all_gather")
        stats["final_blend"] = "master" if shared is None else f"shared by {world} ranks (peer stores into the master's canvas)"
    return res


_PAYLOADS: Dict[tuple, torch.Tensor] = {}


def _payload_buffer(nbytes: int, device) -> torch.Tensor:
    """Reused send buffer of the NCCL transport (stable address: it is baked into the wave graph)."""
    key = (int(nbytes), str(device))
    if key not in _PAYLOADS:
        if len(_PAYLOADS) > 8:
            _PAYLOADS.clear()
        _PAYLOADS[key] = torch.zeros(int(nbytes), dtype=torch.uint8, device=device)
    return _PAYLOADS[key]


def upscale_exact(image: torch.Tensor, denoiser, tile_width: int, tile_height: int, padding: int, mask_blur: int,
                  force_uniform_tiles: bool = True, group=None, stats: Optional[dict] = None) -> torch.Tensor:
    """`semantics="exact"`: N ranks cooperatively execute the SINGLE-GPU
    progressive job, so the result is bit-identical to process_single_gpu at any world size
    (the reference's static mode is not -- SURVEY.md 8c).  Every rank keeps a full canvas;
    the tiles of each dependency wave are split round-robin, each rank crops + samples its
    share, one all-gather per wave exchanges the truncated u8 tiles, and every rank blends
    the whole wave (replicated blend keeps all canvases identical).  Every rank returns the
    result."""
    from . import _native as nat
    from .engine import Canvas, DevicePlan, _require_cuda, _stream_ptr, denoise_packed, _sorted_by_shape
    from .planner import get_plan

    _require_cuda(image, "image")
    rank, world = dist_info(group)
    B, H, W, _ = image.shape
    plan = get_plan(W, H, tile_width, tile_height, padding, mask_blur, force_uniform_tiles)
    with torch.cuda.device(image.device):
        dp = DevicePlan.get(plan, image.device)
        canvas = Canvas(dp, B).load(image)
        n_waves = 0
        for wave in plan.waves():
            wave = _sorted_by_shape(plan, wave)
            shares = [wave[r::world] for r in range(world)]
            mine = shares[rank]
            where, sizes = tile_payload_layout(plan, shares, B)
            payload = torch.zeros(max(sizes[rank], 16), dtype=torch.uint8, device=image.device)
            if mine:
                buf, offs = canvas.crop(mine)
                out = denoise_packed(plan, mine, buf, offs, B, denoiser)
                q = torch.empty(out.numel(), dtype=torch.uint8, device=out.device)
                nat.pack_tiles_u8(out.data_ptr(), q.data_ptr(), out.numel(), _stream_ptr())
                canvas.launches += 1
                for i, tid in enumerate(mine):
                    t = plan.tiles[tid]
                    n = B * t.ph * t.pw * 3
                    payload[where[tid][1]: where[tid][1] + n] = q[int(offs[i]): int(offs[i]) + n]
            gathered, _ = all_gather_bytes(payload, group)
            cap = gathered.shape[1]
            offs_all = np.array([where[t][0] * cap + where[t][1] for t in wave], dtype=np.int64)
            canvas.blend(wave, gathered.view(-1), offs_all)
            n_waves += 1
        res = canvas.result()
    if stats is not None:
        stats["gpu_launches"] = stats.get("gpu_launches", 0) + canvas.launches
        stats["algo_bytes"] = stats.get("algo_bytes", 0) + canvas.algo_bytes
        stats["tiles"], stats["waves"] = len(plan.tiles), n_waves
    return res


# --------------------------------------------------------------------------------------
# collector
# --------------------------------------------------------------------------------------
def collector_order(world: int, enabled_worker_ids: Sequence[str], worker_id_of_rank: Sequence[str]) -> List[int]:
    """Rank order of the collected batch: master (rank 0) first, then workers in the order of
    `enabled_worker_ids` with repeated ids dropped (the node de-duplicates the list before it assembles,
    nodes/collector.py:245-253; the assembly loop itself, :193-223, would repeat them), then
    unexpected ids sorted."""
    order = [0]
    rank_of = {str(w): r for r, w in enumerate(worker_id_of_rank) if r != 0}
    seen = set()
    for w in [str(x) for x in enabled_worker_ids]:
        if w in seen:
            continue
        seen.add(w)
        if w in rank_of:
            order.append(rank_of[w])
    for w in sorted(rank_of):
        if w not in seen:
            order.append(rank_of[w])
    return order


def gather_image_payloads(payload: torch.Tensor, shape: Sequence[int], group=None):
    """All-gather one u8 image batch per rank; shapes may differ in the batch dimension.
    -> list of u8 tensors [B_r, H, W, C] indexed by rank."""
    rank, world = dist_info(group)
    meta = torch.tensor(list(shape), dtype=torch.int64, device=payload.device)
    metas = torch.empty((world, 4), dtype=torch.int64, device=payload.device)
    if world == 1:
        metas[0] = meta
    else:
        td.all_gather_into_tensor(metas.view(-1), meta, group=group)
    buf, _ = all_gather_bytes(payload.reshape(-1), group)
    out = []
    for r in range(world):
        b, h, w, c = [int(v) for v in metas[r].tolist()]
        out.append(buf[r, : b * h * w * c].view(b, h, w, c))
    return out
