"""mie_aux.Cache (src/pymiecoated/pymiecoated/mie_aux.py:22-33): FIFO dictionary of bounded size."""


class Cache(dict):
    def __init__(self, size=10):
        super().__init__()
        self.size = size
        self.log = []

    def __setitem__(self, key, value):
        super().__setitem__(key, value)
        self.log.append(key)
        if len(self.log) > self.size:
            del self[self.log[0]]
            self.log.pop(0)
