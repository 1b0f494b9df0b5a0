"""Byte-budgeted LRU for device-resident caches (decoded seed volumes, segmentations, packed subjects)."""
from __future__ import annotations

from collections import OrderedDict


class ByteLRU:
    """``get`` / ``put`` with least-recently-used eviction once the stored bytes exceed ``budget``.  The newest
    entry is never evicted, so a single over-sized item still works.  ``budget=None``: unbounded."""

    def __init__(self, budget: int | None):
        self.budget, self.bytes = budget, 0
        self._d: OrderedDict = OrderedDict()

    def get(self, key):
        hit = self._d.get(key)
        if hit is None:
            return None
        self._d.move_to_end(key)
        return hit[0]

    def put(self, key, value, nbytes: int):
        old = self._d.pop(key, None)
        if old is not None:
            self.bytes -= old[1]
        self._d[key] = (value, int(nbytes))
        self.bytes += int(nbytes)
        while self.budget is not None and self.bytes > self.budget and len(self._d) > 1:
            _, (_, nb) = self._d.popitem(last=False)
            self.bytes -= nb

    def __setitem__(self, key, value):
        """Dictionary-style insert; the size is taken from the tensor / array when it has one."""
        nbytes = getattr(value, "nbytes", None)
        if nbytes is None and hasattr(value, "numel"):
            nbytes = value.numel() * value.element_size()
        self.put(key, value, int(nbytes or 0))

    def __len__(self):
        return len(self._d)

    def __contains__(self, key):
        return key in self._d
