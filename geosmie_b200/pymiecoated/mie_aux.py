"""mie_aux.Cache (src/pymiecoated/pymiecoated/mie_aux.py:22-33): dictionary that keeps the `size` most recently inserted keys."""
from collections import deque


class Cache(dict):
    """Same behaviour as the reference's class: every assignment counts as an insertion (also of a key that is already present) and
    the oldest insertion is evicted once more than `size` have been made."""

    def __init__(self, size=10):
        dict.__init__(self)
        self.size = size
        self._order = deque()

    def __setitem__(self, key, value):
        dict.__setitem__(self, key, value)
        self._order.append(key)
        while len(self._order) > self.size:
            dict.__delitem__(self, self._order.popleft())
