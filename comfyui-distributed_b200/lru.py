"""Bounded caches of the host side: plans, device-resident tables, captured graphs, symmetric-memory jobs.

A job's geometry selects its plan, tables and graphs; a ComfyUI workflow that alternates between a few geometries
(two upscale nodes in one graph, a batch of differently sized images) must keep all of them warm, so the caches
evict the LEAST RECENTLY USED entry, one at a time, instead of dropping everything when they fill.  Every rank of
a multi-GPU job looks its entries up in the same order, so the ranks evict (and collectively re-create) the same
entries at the same call."""
from __future__ import annotations

from collections import OrderedDict
from typing import Callable, Generic, Hashable, Optional, TypeVar

V = TypeVar("V")


class LruCache(Generic[V]):
    def __init__(self, capacity: int):
        assert capacity >= 1
        self.capacity = int(capacity)
        self._d: "OrderedDict[Hashable, V]" = OrderedDict()

    def get(self, key: Hashable, default: Optional[V] = None) -> Optional[V]:
        if key not in self._d:
            return default
        self._d.move_to_end(key)
        return self._d[key]

    def put(self, key: Hashable, value: V) -> V:
        self._d[key] = value
        self._d.move_to_end(key)
        while len(self._d) > self.capacity:
            self._d.popitem(last=False)
        return value

    def get_or_build(self, key: Hashable, build: Callable[[], V], valid: Optional[Callable[[V], bool]] = None) -> V:
        """The cached value when present (and still `valid`: id()-keyed entries check the object they were built for),
        else build(), stored as the most recent entry."""
        if key in self._d:
            v = self._d[key]
            if valid is None or valid(v):
                self._d.move_to_end(key)
                return v
        return self.put(key, build())

    def __contains__(self, key: Hashable) -> bool:
        return key in self._d

    def __len__(self) -> int:
        return len(self._d)

    def values(self):
        return self._d.values()

    def keys(self):
        return self._d.keys()

    def clear(self) -> None:
        self._d.clear()
