"""ImageBatchDivider (reference nodes/utilities.py:7-20, 235-268): split an
IMAGE batch into up to 10 contiguous, near-equal chunks.  Zero-copy views on whatever device
the batch lives on -- the natural partner of DistributedCollector's gathered batch."""
from __future__ import annotations

MAX_PARTS = 10


def chunk_bounds(total_items: int, n_splits: int):
    """Contiguous [start, end) bounds; the first `total % n` chunks get one extra item."""
    n = max(1, int(n_splits))
    total = max(0, int(total_items))
    base, extra = divmod(total, n)
    out, start = [], 0
    for i in range(n):
        end = start + base + (1 if i < extra else 0)
        out.append((start, end))
        start = end
    return out


class _Wildcard(str):
    """ComfyUI's link validation compares types with `!=`; a type that is never unequal connects to
    anything (the reference's AnyType, nodes/utilities.py:79-83)."""

    def __ne__(self, other) -> bool:
        return False


class _AnyTuple(tuple):
    """ComfyUI indexes RETURN_TYPES per connected output: every index answers with the wildcard type,
    like the reference's ByPassTypeTuple (nodes/utilities.py:226-233); iteration still yields the
    declared entries."""

    def __getitem__(self, index):
        item = super().__getitem__(0 if isinstance(index, int) and index > 0 else index)
        return _Wildcard("*") if isinstance(item, str) else item


class ImageBatchDivider:
    @classmethod
    def INPUT_TYPES(s):
        return {"required": {
            "images": ("IMAGE",),
            "divide_by": ("INT", {"default": 2, "min": 1, "max": MAX_PARTS, "step": 1, "display": "number",
                                  "tooltip": "Number of parts to divide the batch into"}),
        }}

    RETURN_TYPES = _AnyTuple(("IMAGE",))
    RETURN_NAMES = _AnyTuple(tuple(f"batch_{i + 1}" for i in range(MAX_PARTS)))
    FUNCTION = "divide_batch"
    OUTPUT_NODE = True
    CATEGORY = "image"

    def divide_batch(self, images, divide_by):
        parts = max(1, min(int(divide_by), MAX_PARTS))
        empty = images[:0]
        outs = [images[a:b] if b > a else empty for a, b in chunk_bounds(images.shape[0], parts)]
        outs += [empty] * (MAX_PARTS - len(outs))
        return tuple(outs[:MAX_PARTS])
