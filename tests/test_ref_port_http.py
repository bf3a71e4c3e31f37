"""The HTTP + PNG static-mode port (the --impl reference arm for N > 1) really runs over
aiohttp on 127.0.0.1, and its result equals oracle.replay_static for the pull order the
master observed (the same check the survey did against the real reference)."""
import numpy as np
import torch

import ref_port_http
import usdu_oracle as orc


def test_http_static_mode_equals_replay():
    workload = (1, 300, 420, 128, 16, 8)
    B, H, W, tile, pad, blur = workload
    T = len(orc.calculate_tiles(W, H, tile, tile))
    r = ref_port_http.run_job(workload, 9, 0.5, participants=3, tile_ids=list(range(T)))
    assert sorted(r["master_tiles"] + r["worker_tiles"]) == list(range(T))
    asg = [r["master_tiles"]] + [r["pulls"].get(w, []) for w in ("w1", "w2")]
    assert sorted(t for a in asg for t in a) == list(range(T))
    g = torch.Generator().manual_seed(0)
    img = (torch.floor(torch.rand(B, H, W, 3, generator=g) * 255) / 255).numpy()
    ref = orc.replay_static(img, orc.make_t0_denoiser(9, 0.5), tile, tile, pad, blur, True, asg)
    assert np.array_equal(r["output"].numpy(), ref)
    assert sum(w["tiles"] for w in r["workers"]) == len(r["worker_tiles"])
