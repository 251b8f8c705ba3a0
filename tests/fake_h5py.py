"""An in-memory stand-in for `h5py.File` (h5py is absent from this image): a dict of NumPy arrays per path with the access
patterns the save/restore code uses - `f[name] = value` creates a dataset, `f[name][index] = value` writes into it, modes
"w" (truncate) and "a" (append), context manager.  Lets the tests pin the on-disk contract of the reference - dataset
names, shapes and dtypes of qoc/models/schroedingermodels.py:258-313 and qoc/models/lindbladmodels.py:254-339.
Test infrastructure only."""
import numpy as np

FILES = {}


class File(object):
    def __init__(self, path, mode="r"):
        if mode == "w":
            FILES[path] = {}
        elif path not in FILES:
            if mode == "a":
                FILES[path] = {}
            else:
                raise OSError("no such file: {}".format(path))
        self._d = FILES[path]

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    def __setitem__(self, name, value):
        if name in self._d:
            raise ValueError("Unable to create dataset (name already exists)")          # h5py's behaviour
        self._d[name] = np.array(value)

    def __getitem__(self, name):
        return self._d[name]

    def __contains__(self, name):
        return name in self._d

    def keys(self):
        return self._d.keys()


class _Lock(object):
    def __init__(self, path):
        self.path = path

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


class Timeout(Exception):
    pass


def install(monkeypatch):
    """patch qoc_b200.models.state to write through this module; returns the FILES dict (cleared)."""
    import sys
    import types
    from qoc_b200.models import state
    FILES.clear()
    mod = types.ModuleType("h5py")
    mod.File = File
    monkeypatch.setattr(state, "h5py", mod)
    monkeypatch.setattr(state, "FileLock", _Lock)
    monkeypatch.setattr(state, "Timeout", Timeout)
    monkeypatch.setitem(sys.modules, "h5py", mod)
    return FILES
