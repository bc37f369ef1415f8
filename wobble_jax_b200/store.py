"""Persistence container for ``model.save`` / ``model.load`` / ``Data.save`` / ``Data.load``.

The reference writes HDF5 through h5py (jabble/model.py:35-82, 576-593, 903-909, 1030-1037,
1145-1157; jabble/dataset.py:239-277).  h5py is not a dependency of this package: when it can be
imported, files are real HDF5 with the reference's group / dataset names; otherwise the same tree
(same names, same order, same attributes) is kept in memory and written as one ``.npz`` archive
whose keys are the HDF5 paths.  Either way ``save`` / ``load`` code is written once, against the
small h5py-like interface below.
"""

import json
import zipfile

import numpy as np

try:                                   # pragma: no cover - not installed in the build image
    import h5py
except ImportError:                    # the .npz mirror
    h5py = None

_ORDER = "__order__"
_ATTRS = "__attrs__"


class Dataset:
    """A leaf: ``ds[()]``, ``ds[:]``, ``.dtype``, ``.asstr()`` like an h5py dataset."""

    def __init__(self, name, value):
        self.name = name
        self.value = np.asarray(value)

    @property
    def dtype(self):
        return self.value.dtype

    def __getitem__(self, idx):
        if isinstance(idx, tuple) and idx == ():
            return self.value[()]
        return self.value[idx]

    def __array__(self, dtype=None, copy=None):
        return self.value if dtype is None else self.value.astype(dtype)

    def asstr(self):
        v = self.value
        if v.dtype.kind == "S":
            v = np.char.decode(v, "utf-8")
        return v.astype(str)


class Group:
    """An ordered group: ``create_group``, ``create_dataset``, ``keys()``, ``[]``, ``.name``, ``.attrs``."""

    def __init__(self, name="/"):
        self.name = name
        self.items = {}
        self.attrs = {}

    def _child(self, key):
        return (self.name.rstrip("/") + "/" + key) if self.name != "/" else "/" + key

    def create_group(self, key, track_order=None):
        if key in self.items:
            raise ValueError(f"{key} already exists in {self.name}")
        g = Group(self._child(key))
        self.items[key] = g
        return g

    def create_dataset(self, key, data=None, dtype=None):
        if key in self.items:
            raise ValueError(f"{key} already exists in {self.name}")
        if isinstance(data, str):
            data = np.bytes_(data.encode("utf-8"))       # h5py string_dtype scalars read back as bytes
        self.items[key] = Dataset(self._child(key), np.asarray(data))
        return self.items[key]

    def keys(self):
        return self.items.keys()

    def __contains__(self, key):
        return key in self.items

    def __getitem__(self, key):
        node = self
        for part in key.strip("/").split("/"):
            node = node.items[part]
        return node

    # ---- flat (path -> array) form for the .npz mirror
    def _flatten(self, out, order, attrs):
        order[self.name] = list(self.items.keys())
        if self.attrs:
            attrs[self.name] = {k: (v if isinstance(v, (str, int, float)) else np.asarray(v).tolist())
                                for k, v in self.attrs.items()}
        for key, item in self.items.items():
            if isinstance(item, Group):
                item._flatten(out, order, attrs)
            else:
                out[item.name] = item.value


def is_string_dtype(dtype):
    """h5py.check_string_dtype(...) is not None, for both backends."""
    if h5py is not None and not isinstance(dtype, np.dtype):
        return h5py.check_string_dtype(dtype) is not None
    if h5py is not None and h5py.check_string_dtype(dtype) is not None:
        return True
    return np.dtype(dtype).kind in ("S", "U", "O")


def string_dataset(group, key, text):
    """group.create_dataset(key, data=text, dtype=h5py.string_dtype('utf-8'))."""
    if h5py is not None and not isinstance(group, Group):
        return group.create_dataset(key, data=text, dtype=h5py.string_dtype(encoding="utf-8"))
    return group.create_dataset(key, data=text)


class File:
    """``with File(filename, 'w') as hf`` -- an h5py.File when h5py is there, else the mirror."""

    def __init__(self, filename, mode="r"):
        self.filename, self.mode = filename, mode
        self.h5 = None
        self.root = None

    def __enter__(self):
        if self.mode == "r":
            if zipfile.is_zipfile(self.filename):
                self.root = _read_npz(self.filename)
                return self.root
            if h5py is None:
                raise OSError(f"{self.filename} is not a .npz mirror and h5py is not installed")
            self.h5 = h5py.File(self.filename, "r")
            return self.h5
        if h5py is not None:
            self.h5 = h5py.File(self.filename, "w")
            return self.h5
        self.root = Group("/")
        return self.root

    def __exit__(self, exc_type, exc, tb):
        if self.h5 is not None:
            self.h5.close()
        elif self.mode != "r" and exc_type is None:
            out, order, attrs = {}, {}, {}
            self.root._flatten(out, order, attrs)
            out[_ORDER] = np.array(json.dumps(order))
            out[_ATTRS] = np.array(json.dumps(attrs))
            with open(self.filename, "wb") as f:      # exactly the name the caller gave (no ".npz" appended)
                np.savez(f, **out)
        return False


def _read_npz(filename):
    with np.load(filename, allow_pickle=False) as z:
        order = json.loads(str(z[_ORDER]))
        attrs = json.loads(str(z[_ATTRS]))
        arrays = {k: z[k] for k in z.files if k not in (_ORDER, _ATTRS)}

    def build(name):
        g = Group(name)
        g.attrs = dict(attrs.get(name, {}))
        for key in order[name]:
            child = g._child(key)
            if child in order:
                g.items[key] = build(child)
            else:
                g.items[key] = Dataset(child, arrays[child])
        return g

    return build("/")
