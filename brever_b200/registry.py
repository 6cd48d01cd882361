"""Name -> object registry with the reference's contract (brever/registry.py:1-23):
duplicate registration raises ValueError, unknown lookup raises KeyError."""


class Registry:
    def __init__(self, tag):
        self.tag = tag
        self._entries = {}

    def register(self, name):
        def decorator(obj):
            if name in self._entries:
                raise ValueError(f'"{name}" already in {self.tag} registry')
            self._entries[name] = obj
            return obj
        return decorator

    def get(self, name):
        try:
            return self._entries[name]
        except KeyError:
            raise KeyError(f'"{name}" not in {self.tag} registry') from None

    def keys(self):
        return self._entries.keys()
