from . import makedirs  # noqa: F401


class Data:
    """Attribute bag standing in for ``torch_geometric.data.Data`` (what datasets_3D.py:24-67 touches: attribute
    get/set, ``in``, ``keys``).  TEST INFRASTRUCTURE."""

    def __init__(self, **kw):
        self.__dict__.update(kw)

    def __contains__(self, key):
        return key in self.__dict__

    def __getitem__(self, key):
        return self.__dict__[key]

    def __setitem__(self, key, value):
        self.__dict__[key] = value

    @property
    def keys(self):
        return [k for k in self.__dict__ if not k.startswith("__")]


class InMemoryDataset:
    """Base-class stand-in: the fixtures call ``subgraph`` on an instance made with ``__new__`` (no disk access)."""

    def __init__(self, *a, **k):
        pass
