"""oracle.collector_combine == the reference's own DistributedCollector arithmetic (worker PNG round trip,
master decode, _reorder_and_combine_tensors), loaded from /root/reference by oracle/ref_collector.py.
Skipped where the reference tree is absent; tests/golden/collector_ref.json pins the same there."""
import hashlib
import json
import os

import numpy as np
import pytest
import torch

import ref_collector
import usdu_oracle as orc

G = os.path.join(os.path.dirname(__file__), "golden")
CASES = json.load(open(os.path.join(G, "collector_ref.json")))["cases"]


def _inputs(case):
    g = torch.Generator().manual_seed(case["seed"])
    H, W = case["H"], case["W"]
    master = torch.rand(case["master_b"], H, W, 3, generator=g)
    workers = {w: torch.rand(b, H, W, 3, generator=g) for w, b in case["workers"]}
    return master, workers


@pytest.mark.parametrize("case", CASES, ids=lambda c: c["name"])
def test_oracle_collector_matches_golden(case):
    master, workers = _inputs(case)
    out = orc.collector_combine(master.numpy(), {w: t.numpy() for w, t in workers.items()}, case["order"], case["delegate"])
    assert hashlib.sha256(np.ascontiguousarray(out, dtype=np.float32).tobytes()).hexdigest() == case["sha256"]
    assert list(out.shape) == case["shape"]


@pytest.mark.skipif(not ref_collector.available(), reason="reference tree not present")
@pytest.mark.parametrize("case", CASES, ids=lambda c: c["name"])
def test_reference_collector_matches_golden_and_oracle(case):
    master, workers = _inputs(case)
    ref = ref_collector.combine(master, workers, case["order"], case["delegate"]).numpy()
    assert hashlib.sha256(np.ascontiguousarray(ref, dtype=np.float32).tobytes()).hexdigest() == case["sha256"]
    out = orc.collector_combine(master.numpy(), {w: t.numpy() for w, t in workers.items()}, case["order"], case["delegate"])
    assert np.array_equal(out, ref)
