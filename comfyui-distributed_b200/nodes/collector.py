"""DistributedCollector -- same node signature as the reference's nodes/collector.py:24-56.

The reference's workers PNG+base64-encode every image and POST it to the master
(collector.py:84-119 -> api/job_routes.py:273-343); here every rank contributes its batch
to one NCCL all-gather of u8 images and rank 0 assembles the result in the reference's
order: master's images first (kept at full fp32 precision, collector.py:276), then each
enabled worker's images (which went through the truncating u8 cast, collector.py:95-98),
then unexpected workers sorted by id (collector.py:193-236).  Audio stays in Python.
"""
from __future__ import annotations

import json

import torch

from .. import dist as usdu_dist


def _native_pack(images: torch.Tensor) -> torch.Tensor:
    """fp32 [B,H,W,C] (any device) -> u8 CUDA tensor, trunc(255*x) on the GPU."""
    from .. import _native as nat
    dev = images.device if images.is_cuda else torch.device("cuda", torch.cuda.current_device())
    x = images.to(torch.float32).contiguous()
    if not x.is_cuda:
        x = (x if x.is_pinned() else x.pin_memory()).to(dev, non_blocking=True)
    q = torch.empty(x.shape, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        nat.pack_tiles_u8(x.data_ptr(), q.data_ptr(), x.numel(), torch.cuda.current_stream().cuda_stream)
    return q


def _native_unpack(q: torch.Tensor) -> torch.Tensor:
    from .. import _native as nat
    out = torch.empty(q.shape, dtype=torch.float32, device=q.device)
    q = q.contiguous()
    with torch.cuda.device(q.device):
        nat.unpack_tiles_f32(q.data_ptr(), out.data_ptr(), q.numel(), torch.cuda.current_stream().cuda_stream)
    return out


def collect_images(images: torch.Tensor, enabled_worker_ids, worker_id: str, delegate_only: bool = False,
                   group=None, pack=_native_pack, unpack=_native_unpack, audio=None):
    """SPMD collective behind the node: the u8 batches and the small per-rank records (worker id, audio) travel to
    rank 0 only (dist.gather_to_root).  Returns (combined CPU batch, rank order, audio pieces by rank) on rank 0 and
    (None, None, None) on the other ranks."""
    rank, world = usdu_dist.dist_info(group)
    q = pack(images)
    parts, records = usdu_dist.gather_to_root(q, {"worker_id": str(worker_id), "audio": audio}, group)
    if rank != 0:
        return None, None, None
    ids = [rec["worker_id"] for rec in records]
    order = usdu_dist.collector_order(world, enabled_worker_ids, ids)
    out = []
    for r in order:
        if r == 0:
            if not delegate_only:
                out.append(images.detach().to("cpu", torch.float32).contiguous())
        elif parts[r].numel() > 0:
            out.append(unpack(parts[r]).cpu())
    if not out:
        raise ValueError("No image data collected from master or workers")
    return torch.cat(out, dim=0).contiguous(), order, [rec["audio"] for rec in records]


def combine_audio(pieces, empty_audio):
    """collector.py:121-174 -- concatenate waveforms along the sample axis, master first."""
    waves, rate = [], 44100
    for a in pieces:
        if a is None:
            continue
        w = a.get("waveform")
        if w is not None and w.numel() > 0:
            if not waves or rate == 44100:
                rate = a.get("sample_rate", 44100)
            waves.append(w)
    if not waves:
        return empty_audio
    try:
        return {"waveform": torch.cat(waves, dim=-1), "sample_rate": rate}
    except Exception:
        return empty_audio


class DistributedCollectorNode:
    EMPTY_AUDIO = {"waveform": torch.zeros(1, 2, 1), "sample_rate": 44100}

    @classmethod
    def INPUT_TYPES(s):
        return {
            "required": {
                "images": ("IMAGE",),
                "load_balance": ("BOOLEAN", {
                    "default": False,
                    "tooltip": "Run this workflow on one least-busy participant (master included when participating).",
                }),
            },
            "optional": {"audio": ("AUDIO",)},
            "hidden": {
                "multi_job_id": ("STRING", {"default": ""}),
                "is_worker": ("BOOLEAN", {"default": False}),
                "master_url": ("STRING", {"default": ""}),
                "enabled_worker_ids": ("STRING", {"default": "[]"}),
                "worker_batch_size": ("INT", {"default": 1, "min": 1, "max": 1024}),
                "worker_id": ("STRING", {"default": ""}),
                "pass_through": ("BOOLEAN", {"default": False}),
                "delegate_only": ("BOOLEAN", {"default": False}),
            },
        }

    RETURN_TYPES = ("IMAGE", "AUDIO")
    RETURN_NAMES = ("images", "audio")
    FUNCTION = "run"
    CATEGORY = "image"

    def run(self, images, load_balance=False, audio=None, multi_job_id="", is_worker=False, master_url="",
            enabled_worker_ids="[]", worker_batch_size=1, worker_id="", pass_through=False, delegate_only=False):
        empty_audio = {"waveform": torch.zeros(1, 2, 1), "sample_rate": 44100}
        if not multi_job_id or pass_through:
            return (images, audio if audio is not None else empty_audio)
        rank, world = usdu_dist.dist_info()
        enabled = [str(w) for w in json.loads(enabled_worker_ids)]
        if world == 1:  # no participants besides the master (collector.py:255-256)
            if is_worker:
                # started as an HTTP worker of the reference's orchestrator: there is no torch.distributed peer to send the
                # batch to, and the reference's HTTP collection is not part of this package -- say so instead of dropping it
                import warnings
                warnings.warn("DistributedCollector (B200): running as an HTTP worker of the reference's orchestrator is not "
                              "supported (launch one rank per GPU with torch.distributed); this worker's images stay local.",
                              RuntimeWarning, stacklevel=2)
            return (images, audio if audio is not None else empty_audio)
        wid = worker_id if (worker_id or rank == 0) else f"rank{rank}"
        if not enabled:  # SPMD launch without the reference's orchestrator: every rank is enabled
            enabled = [f"rank{r}" for r in range(1, world)]
        combined, order, audios = collect_images(images, enabled, wid, delegate_only=delegate_only, audio=audio)
        if rank != 0:
            return (images, audio if audio is not None else self.EMPTY_AUDIO)
        pieces = [audios[r] for r in order if not (r == 0 and delegate_only)]
        return (combined, combine_audio(pieces, empty_audio))
