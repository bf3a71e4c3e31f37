"""Drop-in boundary: our node classes present ComfyUI with exactly the reference's signatures --
INPUT_TYPES (names, order, types, widget options), RETURN_TYPES / RETURN_NAMES / FUNCTION / CATEGORY /
OUTPUT_NODE, the entry method's parameters and defaults, IS_CHANGED, and the registration keys.
tests/golden/node_signatures.json is dumped from the REAL node classes by oracle/ref_signatures.py."""
import json
import os
import sys

import pytest

from __graft_entry__ import load_package

load_package()
import comfyui_distributed_b200 as pkg  # noqa: E402
from comfyui_distributed_b200 import nodes as our_nodes  # noqa: E402

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import ref_signatures  # noqa: E402

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "node_signatures.json")))["nodes"]
OURS = {"UltimateSDUpscaleDistributed": our_nodes.UltimateSDUpscaleDistributed,
        "DistributedCollector": our_nodes.DistributedCollectorNode,
        "ImageBatchDivider": our_nodes.ImageBatchDivider}


@pytest.mark.parametrize("name", sorted(OURS))
def test_node_signature_equals_reference(name):
    got, want = ref_signatures.describe(OURS[name]), GOLD[name]
    if name == "UltimateSDUpscaleDistributed":      # ComfyUI's own lists: position and kind only
        for key in ("sampler_name", "scheduler"):
            assert isinstance(got["input_types"]["required"][key][0], list)
            got["input_types"]["required"][key] = want["input_types"]["required"][key]
    assert got["input_order"] == want["input_order"]                 # widget order is positional in saved workflows
    assert json.loads(json.dumps(got["input_types"])) == want["input_types"]
    for key in ("return_types", "return_types_beyond_end", "return_names", "function", "category", "output_node",
                "params", "is_changed_nan"):
        assert json.loads(json.dumps(got[key])) == want[key], key


def test_registration_keys():
    m = GOLD["upscale_mappings"]
    assert set(m["NODE_CLASS_MAPPINGS"]) <= set(pkg.NODE_CLASS_MAPPINGS)
    for k, v in m["NODE_DISPLAY_NAME_MAPPINGS"].items():
        assert pkg.NODE_DISPLAY_NAME_MAPPINGS[k] == v
    assert {"DistributedCollector", "ImageBatchDivider"} <= set(pkg.NODE_CLASS_MAPPINGS)


def test_divider_outputs_are_wildcards_like_the_reference():
    rt = our_nodes.ImageBatchDivider.RETURN_TYPES
    assert not (rt[0] != "IMAGE") and not (rt[7] != "MASK") and rt[3] == "*"      # never unequal to any type
    assert tuple(rt) == ("IMAGE",) and len(tuple(our_nodes.ImageBatchDivider.RETURN_NAMES)) == 10


def test_error_behaviour_equals_reference():
    """Same exception types (and the same message for the batch rule) as the reference's run(), raised
    before any device work (so this runs without a GPU)."""
    import torch
    from comfyui_distributed_b200.testing import T0Model
    want = GOLD["errors"]
    node = our_nodes.UltimateSDUpscaleDistributed()
    base = (T0Model(), None, None, None, 0, 20, 8.0, "euler", "normal", 0.5, 64, 64, 8, 8, True, False)
    with pytest.raises(ValueError) as e:
        node.run(torch.zeros(2, 64, 64, 3), *base)
    assert [type(e.value).__name__, str(e.value)] == want["batch_of_2_master"]
    with pytest.raises(json.JSONDecodeError) as e:
        node.run(torch.zeros(1, 64, 64, 3), *base, multi_job_id="j", is_worker=False, enabled_worker_ids="[not json")
    assert type(e.value).__name__ == want["malformed_enabled_worker_ids"][0]
    # the batch rule is the master's only (nodes/distributed_upscale.py:137): a worker with B = 2 gets past it
    try:
        node.run(torch.zeros(2, 64, 64, 3), *base, is_worker=True)
    except ValueError as err:                      # pragma: no cover
        assert "4n+1" not in str(err)
    except Exception:                              # no GPU here: NativeError / CUDA error further down is fine
        pass
