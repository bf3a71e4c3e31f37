"""Dump the ComfyUI-facing signatures of the reference's nodes on the hot path (TEST INFRASTRUCTURE, build
container only): INPUT_TYPES(), RETURN_TYPES / RETURN_NAMES / FUNCTION / CATEGORY / OUTPUT_NODE, the entry
method's parameters with defaults, and the registration mappings -- loaded from the REAL files under
/root/reference (nodes/distributed_upscale.py, nodes/collector.py, nodes/utilities.py).
`python oracle/ref_signatures.py` writes tests/golden/node_signatures.json."""
from __future__ import annotations

import inspect
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)


def describe(cls) -> dict:
    fn = getattr(cls, cls.FUNCTION)
    params = []
    for name, p in inspect.signature(fn).parameters.items():
        if name == "self":
            continue
        params.append([name, None if p.default is inspect.Parameter.empty else p.default,
                       p.default is not inspect.Parameter.empty])
    it = cls.INPUT_TYPES()

    def norm(v):            # tuples -> lists so that JSON round-trips compare equal
        if isinstance(v, (tuple, list)):
            return [norm(x) for x in v]
        if isinstance(v, dict):
            return {k: norm(x) for k, x in v.items()}
        return v

    rt = cls.RETURN_TYPES

    def beyond(t):          # the divider's RETURN_TYPES answers any output index (ByPassTypeTuple)
        try:
            return t[len(tuple(t)) + 6]
        except IndexError:
            return None

    return {"input_types": norm(it), "input_order": {k: list(v) for k, v in it.items()},
            "return_types": [rt[i] for i in range(len(tuple(rt)))], "return_types_beyond_end": beyond(rt),
            "return_names": list(getattr(cls, "RETURN_NAMES", ()) or ()), "function": cls.FUNCTION, "category": cls.CATEGORY,
            "output_node": bool(getattr(cls, "OUTPUT_NODE", False)), "params": params,
            "is_changed_nan": (lambda v: v != v)(cls.IS_CHANGED()) if hasattr(cls, "IS_CHANGED") else None}


def reference_signatures() -> dict:
    import ref_collector
    import ref_static_run
    env = ref_static_run._Env()
    try:
        up_mod = env.mods["nodes.distributed_upscale"]
        up = describe(up_mod.UltimateSDUpscaleDistributed)
        # the sampler / scheduler lists come from ComfyUI (stubbed here): compare their position only
        up["input_types"]["required"]["sampler_name"] = ["<comfy.samplers.KSampler.SAMPLERS>"]
        up["input_types"]["required"]["scheduler"] = ["<comfy.samplers.KSampler.SCHEDULERS>"]
        import torch
        errors = {}
        node = up_mod.UltimateSDUpscaleDistributed()
        base = (None, None, None, None, 0, 20, 8.0, "euler", "normal", 0.5, 64, 64, 8, 8, True, False)
        try:
            node.run(torch.zeros(2, 64, 64, 3), *base)
        except Exception as e:      # noqa: BLE001
            errors["batch_of_2_master"] = [type(e).__name__, str(e)]
        try:
            node.run(torch.zeros(1, 64, 64, 3), *base, multi_job_id="j", is_worker=False, enabled_worker_ids="[not json")
        except Exception as e:      # noqa: BLE001
            errors["malformed_enabled_worker_ids"] = [type(e).__name__]
        out = {"UltimateSDUpscaleDistributed": up, "errors": errors,
               "upscale_mappings": {"NODE_CLASS_MAPPINGS": sorted(up_mod.NODE_CLASS_MAPPINGS),
                                    "NODE_DISPLAY_NAME_MAPPINGS": dict(up_mod.NODE_DISPLAY_NAME_MAPPINGS)}}
    finally:
        env.close()
    collector, _, _ = ref_collector.load()
    out["DistributedCollector"] = describe(collector.DistributedCollectorNode)
    util = ref_collector._load("nodes.utilities", "nodes/utilities.py")
    out["ImageBatchDivider"] = describe(util.ImageBatchDivider)
    return out


if __name__ == "__main__":
    sig = reference_signatures()
    path = os.path.join(os.path.dirname(HERE), "tests", "golden", "node_signatures.json")
    with open(path, "w") as f:
        json.dump({"generator": "oracle/ref_signatures.py", "reference": "a91f9fb", "nodes": sig}, f, indent=1, sort_keys=True)
    print("wrote", path, {k: len(json.dumps(v)) for k, v in sig.items()})
