"""Generate tests/golden/bench_digests.json: SHA-256 of the u8 result of the FULL-SIZE bench workloads.

TEST INFRASTRUCTURE ONLY (see oracle/usdu_oracle.py header); run in the build container:

    python oracle/gen_bench_digests.py real cfg2_4k_to_8k_sdxl_512px      # the REAL reference, ~30 min of CPU
    python oracle/gen_bench_digests.py real cfg5_video_17f_4k
    python oracle/gen_bench_digests.py oracle                             # every oracle-side digest

What is pinned (inputs = bench.py's synthetic canvas, seed 0; sampler = T0, seed 123, denoise 0.5):
  * N = 1, "source": "reference": the reference's own process_single_gpu (upscale/modes/single_gpu.py:8-72)
    loaded from /root/reference by oracle/ref_loader.py, on cfg2 (the headline) and cfg5.
  * N = 1, "source": "oracle": oracle.process_single on cfg2, cfg5 (must equal the reference's digest -- asserted
    when both exist), cfg4 (15360x8640, 2040 tiles) and cfg4alt (8192x8192, 1024 tiles).
  * N = 2/4/8, "source": "oracle": oracle.replay_static (verified against real HTTP runs of the reference,
    tests/golden/static_ref_index.json) with the assignment planner.partition(N) produces, which is stored
    beside the digest (upscale/modes/static.py:521-553 semantics).
bench.py, __graft_entry__.smoke() and tests/test_gpu_fullsize.py compare the CUDA path's result with these.
"""
from __future__ import annotations

import hashlib
import json
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
import usdu_oracle as orc  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "bench_digests.json")
SEED, DENOISE = 123, 0.5
WORKLOADS = {
    # name: (B, H, W, tile, padding, blur)   -- bench.py's table + the 8192^2 alternative of SURVEY.md 8(d)
    "cfg2_4k_to_8k_sdxl_512px": (1, 4320, 7680, 512, 32, 8),
    "cfg1_512_256px": (1, 512, 512, 256, 32, 8),
    "cfg4_16k_256px": (1, 8640, 15360, 256, 32, 8),
    "cfg4alt_8k_256px": (1, 8192, 8192, 256, 32, 8),
    "cfg5_video_17f_4k": (17, 2160, 3840, 512, 32, 8),
}


def canvas(B, H, W) -> np.ndarray:
    g = torch.Generator().manual_seed(0)
    x = torch.rand(B, H, W, 3, generator=g)
    return (torch.floor(x * 255) / 255).numpy()


def sha_u8(res_f32: np.ndarray) -> str:
    out = np.round(res_f32 * 255).astype(np.uint8)
    assert np.array_equal(out.astype(np.float32) / np.float32(255), res_f32)
    return hashlib.sha256(out.tobytes()).hexdigest()


def load_db() -> dict:
    if os.path.isfile(OUT):
        return json.load(open(OUT))
    return {"generator": "oracle/gen_bench_digests.py", "reference": "a91f9fb",
            "input": "torch.rand(B,H,W,3, manual_seed(0)) floored to k/255", "sampler": f"T0 seed {SEED} denoise {DENOISE}",
            "digests": {}}


def store(key: str, entry: dict):
    import fcntl
    with open(OUT + ".lock", "w") as lk:             # several generator processes may run side by side
        fcntl.flock(lk, fcntl.LOCK_EX)
        db = load_db()
        db["digests"][key] = entry
        with open(OUT, "w") as f:
            json.dump(db, f, indent=1, sort_keys=True)
    print("stored", key, entry["sha256"][:16], flush=True)


def gen_real(name: str):
    import ref_loader
    from gen_golden import torch_t0
    B, H, W, tile, pad, blur = WORKLOADS[name]
    node, fake_nodes = ref_loader.make_reference_node()
    fake_nodes.fn = torch_t0()
    img = torch.from_numpy(canvas(B, H, W))
    t0 = time.time()
    (res,) = node.process_single_gpu(img, None, [[torch.zeros(1, 77, 8), {}]], [[torch.zeros(1, 77, 8), {}]], None,
                                     SEED, 20, 8.0, "euler", "normal", DENOISE, tile, tile, pad, blur, True, False)
    dt = time.time() - t0
    store(f"{name}/n1/reference", {"sha256": sha_u8(res.numpy()), "source": "reference",
                                   "how": "process_single_gpu of /root/reference via oracle/ref_loader.py",
                                   "cpu_seconds_wall": round(dt, 1), "torch_threads": torch.get_num_threads()})


def gen_oracle(names):
    from importlib import import_module
    sys.path.insert(0, ROOT)
    from __graft_entry__ import load_package
    load_package()
    planner = import_module("comfyui_distributed_b200.planner")
    den = orc.make_t0_denoiser(SEED, DENOISE)
    for name in names:
        B, H, W, tile, pad, blur = WORKLOADS[name]
        img = canvas(B, H, W)
        t0 = time.time()
        res = orc.process_single(img, den, tile, tile, pad, blur, True)
        store(f"{name}/n1/oracle", {"sha256": sha_u8(res), "source": "oracle", "how": "oracle.process_single",
                                    "cpu_seconds_wall": round(time.time() - t0, 1)})
        if name != "cfg2_4k_to_8k_sdxl_512px":
            continue
        plan = planner.get_plan(W, H, tile, tile, pad, blur, True)
        for n in (2, 4, 8):
            asg = [list(map(int, a)) for a in plan.partition(n)]
            t0 = time.time()
            res = orc.replay_static(img, den, tile, tile, pad, blur, True, asg)
            store(f"{name}/n{n}/oracle", {"sha256": sha_u8(res), "source": "oracle",
                                          "how": "oracle.replay_static(planner.partition(n))", "assignment": asg,
                                          "cpu_seconds_wall": round(time.time() - t0, 1)})
    db = load_db()["digests"]
    for name in WORKLOADS:
        a, b = db.get(f"{name}/n1/reference"), db.get(f"{name}/n1/oracle")
        if a and b:
            assert a["sha256"] == b["sha256"], f"{name}: oracle differs from the real reference"
            print(name, "oracle == reference", flush=True)


if __name__ == "__main__":
    if sys.argv[1] == "real":
        gen_real(sys.argv[2])
    else:
        gen_oracle(sys.argv[2:] or list(WORKLOADS))
