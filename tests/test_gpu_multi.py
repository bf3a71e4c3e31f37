"""Multi-GPU (NCCL) parity: static mode across N ranks == oracle.replay_static with the
same fixed partition; collector over NCCL == oracle.collector_combine.  Needs >= 2 GPUs."""
import json
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as td
import torch.multiprocessing as mp

import usdu_oracle as orc
from __graft_entry__ import load_package
from inputs import make_input

pytestmark = pytest.mark.gpu


def _ngpu():
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _entry(rank, world, port, fn, args):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    load_package()
    td.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        fn(rank, world, *args)
    finally:
        td.destroy_process_group()


def _run(fn, world, *args):
    mp.spawn(_entry, args=(world, _free_port(), fn, args), nprocs=world, join=True)


CASES = [("noise", 1, 300, 420, 128, 16, 8), ("noise", 1, 1100, 1300, 256, 32, 8), ("smooth", 5, 136, 168, 64, 16, 8)]


def _w_static(rank, world, case, transport):
    from comfyui_distributed_b200 import dist as udist, planner
    from comfyui_distributed_b200.denoise import T0Denoiser
    udist.USE_PEER_BLEND = transport != "nccl"
    udist.USE_SHARED_FINAL_BLEND = transport == "peer_shared"
    kind, B, H, W, tile, pad, blur = case
    p = planner.get_plan(W, H, tile, tile, pad, blur, True)
    for job in range(3):                                       # later jobs replay the graphs and reuse the payload buffers
        img = make_input(kind, 21 + job, B, H, W)
        x = torch.from_numpy(img).cuda()
        st = {}
        out = udist.upscale_static(x, T0Denoiser(9, 0.5), tile, tile, pad, blur, True, stats=st)
        if transport != "nccl":                                # NVLink boxes: the blend kernel reads the peers' HBM
            assert st["transport"] == "nvlink peer loads", (st["transport"], udist.PeerPayload.last_error)
            assert st["final_blend"].startswith("sharded" if transport == "peer_shared" else "master"), st["final_blend"]
        else:
            assert st["transport"] == "nccl all_gather"
        if rank != 0:
            assert out is x                                    # workers return their input (static.py:314)
            continue
        ref = orc.replay_static(img, orc.make_t0_denoiser(9, 0.5), tile, tile, pad, blur, True, p.partition(world))
        assert np.array_equal(out.cpu().numpy(), ref)
        assert st["tiles_this_rank"] == len(p.partition(world)[0])


@pytest.mark.parametrize("transport", ["peer_shared", "peer", "nccl"])
@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("case", CASES, ids=lambda c: f"{c[0]}_{c[3]}x{c[2]}_b{c[1]}")
def test_static_mode_matches_replay_oracle(world, case, transport):
    if _ngpu() < world:
        pytest.skip(f"needs {world} GPUs")
    _run(_w_static, world, case, transport)


def _w_recorded(rank, world, case):
    import hashlib
    from comfyui_distributed_b200 import dist as udist
    from comfyui_distributed_b200.denoise import T0Denoiser
    img = make_input(case["kind"], case["seed"], case["B"], case["H"], case["W"])
    x = torch.from_numpy(img).cuda()
    out = udist.upscale_static(x, T0Denoiser(case["denoise_seed"], case["denoise"]), case["tile"], case["tile"],
                               case["padding"], case["mask_blur"], case["uniform"], assignment=case["assignment"])
    if rank == 0:
        q = np.round(out.cpu().numpy() * 255).astype(np.uint8)
        assert hashlib.sha256(q.tobytes()).hexdigest() == case["sha256"]


@pytest.mark.parametrize("case", json.load(open(os.path.join(os.path.dirname(__file__), "golden", "static_ref_index.json")))["cases"],
                         ids=lambda c: c["name"])
def test_static_mode_reproduces_the_real_reference_runs(case):
    """The jobs the REAL reference ran over HTTP (tests/golden/static_ref_index.json: recorded pull-order
    assignment + digest of the master's result) on as many ranks as that run had participants."""
    world = len(case["assignment"])
    if _ngpu() < world:
        pytest.skip(f"needs {world} GPUs")
    _run(_w_recorded, world, case)


def _w_exact(rank, world, case):
    from comfyui_distributed_b200 import dist as udist
    from comfyui_distributed_b200.denoise import T0Denoiser
    kind, B, H, W, tile, pad, blur = case
    img = make_input(kind, 21, B, H, W)
    st = {}
    out = udist.upscale_exact(torch.from_numpy(img).cuda(), T0Denoiser(9, 0.5), tile, tile, pad, blur, True, stats=st)
    ref = orc.process_single(img, orc.make_t0_denoiser(9, 0.5), tile, tile, pad, blur, True)
    assert np.array_equal(out.cpu().numpy(), ref)               # == the SINGLE-GPU result, on every rank


@pytest.mark.parametrize("world", [2, 4])
@pytest.mark.parametrize("case", CASES[:2], ids=lambda c: f"{c[0]}_{c[3]}x{c[2]}_b{c[1]}")
def test_exact_mode_equals_single_gpu_oracle(world, case):
    if _ngpu() < world:
        pytest.skip(f"needs {world} GPUs")
    _run(_w_exact, world, case)


def _w_all_ranks(rank, world):
    from comfyui_distributed_b200 import dist as udist, planner
    from comfyui_distributed_b200.denoise import T0Denoiser
    img = make_input("noise", 5, 1, 300, 420)
    out = udist.upscale_static(torch.from_numpy(img).cuda(), T0Denoiser(9, 0.5), 128, 128, 16, 8, True, all_ranks_result=True)
    p = planner.get_plan(420, 300, 128, 128, 16, 8, True)
    ref = orc.replay_static(img, orc.make_t0_denoiser(9, 0.5), 128, 128, 16, 8, True, p.partition(world))
    assert np.array_equal(out.cpu().numpy(), ref)               # every rank rebuilds the master's result


def test_static_mode_result_on_all_ranks():
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    _run(_w_all_ranks, 2)


def _w_conflict_free(rank, world):
    """A partition without overlapping windows inside a rank: the ranks crop straight from the fp32 image (no quantised
    working canvas, usdu_tile_crop_resize_f32), skip their local blends, and the final canvas is composited slab by slab."""
    from comfyui_distributed_b200 import dist as udist
    from comfyui_distributed_b200.denoise import T0Denoiser
    B, H, W, tile, pad, blur = 2, 256, 1280, 256, 32, 8
    asg = [[0, 2, 4], [1, 3]]
    for job in range(3):
        img = make_input("noise", 41 + job, B, H, W)
        x = torch.from_numpy(img).cuda()
        st = {}
        out = udist.upscale_static(x, T0Denoiser(9, 0.5), tile, tile, pad, blur, True, assignment=asg, stats=st)
        assert st["conflict_free"] and st["final_blend"].startswith("sharded")
        job_obj = list(udist.StaticJob._cache.values())[-1]
        assert job_obj.from_image, "expected the crop-from-image path (tensor-core plan, W % 4 == 0)"
        if rank == 0:
            ref = orc.replay_static(img, orc.make_t0_denoiser(9, 0.5), tile, tile, pad, blur, True, asg)
            assert np.array_equal(out.cpu().numpy(), ref)


def test_conflict_free_partition_crops_from_the_image():
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    _run(_w_conflict_free, 2)


def _w_host(rank, world, case):
    """dist.upscale_static_host: every rank uploads / downloads only its slab; the result tensor lives in shared
    page-locked memory.  Several jobs in a row: held results must not be overwritten, dropped ones are recycled."""
    from comfyui_distributed_b200 import dist as udist, planner
    from comfyui_distributed_b200.denoise import T0Denoiser
    kind, B, H, W, tile, pad, blur = case
    p = planner.get_plan(W, H, tile, tile, pad, blur, True)
    held = []
    for job in range(4):
        img = make_input(kind, 31 + job, B, H, W)
        x = torch.from_numpy(img.copy())
        x = x.pin_memory() if job % 2 else x                    # pageable and pinned inputs
        st = {}
        out = udist.upscale_static_host(x, T0Denoiser(9, 0.5), tile, tile, pad, blur, True, stats=st)
        assert out is not NotImplemented, udist.PeerPayload.last_error
        assert np.array_equal(x.numpy(), img)                   # the caller's tensor is never written
        if rank != 0:
            assert out is None
            continue
        ref = orc.replay_static(img, orc.make_t0_denoiser(9, 0.5), tile, tile, pad, blur, True, p.partition(world))
        assert not out.is_cuda and out.is_pinned() and np.array_equal(out.numpy(), ref)
        if job < 2:
            held.append((out, ref))                             # a consumer that keeps its results (ComfyUI's cache)
        for o, r in held:
            assert np.array_equal(o.numpy(), r)
        del out


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("case", CASES, ids=lambda c: f"{c[0]}_{c[3]}x{c[2]}_b{c[1]}")
def test_static_mode_on_host_tensors_moves_slabs(world, case):
    if _ngpu() < world:
        pytest.skip(f"needs {world} GPUs")
    _run(_w_host, world, case)


def _w_node(rank, world):
    from comfyui_distributed_b200.nodes import DistributedCollectorNode, UltimateSDUpscaleDistributed
    from comfyui_distributed_b200.testing import T0Model
    from comfyui_distributed_b200 import planner
    img = make_input("noise", 7, 1, 260, 300)
    x = torch.from_numpy(img)                                   # host tensor, like ComfyUI
    node = UltimateSDUpscaleDistributed()
    (out,) = node.run(x, T0Model(), None, None, None, 9, 20, 8.0, "euler", "normal", 0.5, 128, 128, 16, 8, True, False,
                      multi_job_id="job", is_worker=rank != 0, enabled_worker_ids='["w1"]', worker_id="" if rank == 0 else "w1")
    if rank == 0:
        p = planner.get_plan(300, 260, 128, 128, 16, 8, True)
        ref = orc.replay_static(img, orc.make_t0_denoiser(9, 0.5), 128, 128, 16, 8, True, p.partition(world))
        assert not out.is_cuda and np.array_equal(out.numpy(), ref)
    else:
        assert out is x
    # semantics = "exact": the same call gives the SINGLE-GPU result, on host and device tensors
    node.semantics = "exact"
    single = orc.process_single(img, orc.make_t0_denoiser(9, 0.5), 128, 128, 16, 8, True)
    for xx in (x, x.cuda()):
        (out,) = node.run(xx, T0Model(), None, None, None, 9, 20, 8.0, "euler", "normal", 0.5, 128, 128, 16, 8, True, False,
                          multi_job_id="job", is_worker=rank != 0, enabled_worker_ids='["w1"]', worker_id="" if rank == 0 else "w1")
        if rank == 0:
            assert out.is_cuda == xx.is_cuda and np.array_equal(out.cpu().numpy(), single)
        else:
            assert out is xx
    # collector: rank r contributes r+1 images
    g = torch.Generator().manual_seed(50 + rank)
    imgs = torch.rand(1 + rank, 40, 56, 3, generator=g)
    col = DistributedCollectorNode()
    res, audio = col.run(imgs, multi_job_id="job", is_worker=rank != 0, enabled_worker_ids='["w1"]',
                         worker_id="" if rank == 0 else "w1")
    if rank == 0:
        w = torch.rand(2, 40, 56, 3, generator=torch.Generator().manual_seed(51))
        ref = orc.collector_combine(imgs.numpy(), {"w1": w.numpy()}, ["w1"])
        assert np.array_equal(res.numpy(), ref)
    else:
        assert res is imgs
    assert audio["sample_rate"] == 44100


def test_nodes_two_ranks():
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    _run(_w_node, 2)
