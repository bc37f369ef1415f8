"""Host-side data containers: ragged per-row spectra -> padded (E, Pmax) blocks.

Mirror of the reference's ``jabble/dataset.py:111-237`` (``DataBlock``, ``Data``, ``DataFrame``,
``blockify``).  Arrays are numpy float64 on the host; the device copy is made once per plan by the
C-ABI library (``jb_plan_create``), which replaces ``jax.device_put``.  HDF5 I/O and continuum
fitting (dataset.py:19-109, 239-277) are out of scope.
"""

import numpy as np


class DataBlock:
    """dataset.py:111-137 -- attribute bag of equally-long arrays, hashed by identity."""

    def __init__(self, **kwargs):
        self.__dict__.update(kwargs)

    def __hash__(self):
        return id(self)

    def __eq__(self, other):
        return self is other

    def __getitem__(self, key):
        return self.__dict__[key]

    def __len__(self):
        return self.__dict__[list(self.__dict__.keys())[0]].shape[0]

    def keys(self):
        return self.__dict__.keys()

    def slice(self, start, stop):
        return DataBlock(**{k: v[start:stop] for k, v in self.__dict__.items()})

    def ele(self, i):
        return DataBlock(**{k: v[i] for k, v in self.__dict__.items()})

    def to_device(self, device):
        """No-op: device residency is owned by the plan (one upload per plan)."""
        return self


class DataFrame:
    """dataset.py:280-291 -- one row (epoch x order) of a spectrum."""

    def __init__(self, xs, ys, yivar, mask):
        self.xs = np.asarray(xs, dtype=np.float64)
        self.ys = np.asarray(ys, dtype=np.float64)
        self.yivar = np.asarray(yivar, dtype=np.float64)
        self.mask = np.asarray(mask).astype(bool)

    def to_device(self, device):
        return self


class Data:
    """dataset.py:147-237."""

    def __init__(self, frames, *args):
        self.dataframes = frames
        self.metadata = {}
        self.metakeys = {}

    def save(self, filename):
        """dataset.py:239-254 -- group ``data`` with one ``dataframe_<i>`` per row, group ``metadata``
        (HDF5 when h5py is installed, else the .npz mirror of the same tree)."""
        from . import store
        with store.File(filename, "w") as hf:
            group = hf.create_group("data")
            for i, dataframe in enumerate(self.dataframes):
                subgroup = group.create_group(f"dataframe_{i}", track_order=True)
                subgroup.create_dataset("xs", data=dataframe.xs)
                subgroup.create_dataset("ys", data=dataframe.ys)
                subgroup.create_dataset("yivar", data=dataframe.yivar)
                subgroup.create_dataset("mask", data=dataframe.mask)
            metagroup = hf.create_group("metadata")
            for key in self.metadata.keys():
                value = np.asarray(self.metadata[key])
                if value.dtype.kind in {"U", "S"}:
                    metagroup.create_dataset(key, data=np.char.encode(value, "utf-8"))
                else:
                    metagroup.create_dataset(key, data=value)

    @staticmethod
    def load(filename):
        """dataset.py:256-277."""
        from . import store
        with store.File(filename, "r") as hf:
            group = hf["data"]
            dataframes = []
            for key in group.keys():
                if key.startswith("dataframe_"):
                    dataframes.append(DataFrame(group[key]["xs"][:], group[key]["ys"][:], group[key]["yivar"][:],
                                                group[key]["mask"][:]))
            data = Data(dataframes)
            metagroup = hf["metadata"]
            for key in metagroup.keys():
                if store.is_string_dtype(metagroup[key].dtype):
                    data.metadata[key] = np.array(metagroup[key].asstr()[:], dtype=str)
                else:
                    data.metadata[key] = np.array(metagroup[key][:])
        return data

    def __getitem__(self, i):
        return self.dataframes[i]

    def __len__(self):
        return len(self.dataframes)

    @property
    def xs(self):
        return [f.xs for f in self.dataframes]

    @property
    def ys(self):
        return [f.ys for f in self.dataframes]

    @property
    def yivar(self):
        return [f.yivar for f in self.dataframes]

    @property
    def mask(self):
        return [f.mask for f in self.dataframes]

    @property
    def yerr(self):
        return [1 / np.sqrt(f.yivar) for f in self.dataframes]

    @staticmethod
    def from_lists(xs, ys, yivar, ma):
        """dataset.py:183-187."""
        return Data([DataFrame(xs[i], ys[i], yivar[i], ma[i]) for i in range(len(xs))])

    def to_device(self, device):
        return self

    def blockify(self, device=None):
        """dataset.py:192-237 -- pad every row to the longest one (x = y = ivar = 0, mask = True)
        and turn every metadata key into dense integer ids.  Returns (datablock, metablock)."""
        # The same frames give the same block OBJECTS, so that everything downstream keyed on the block
        # (the resident copy in HBM: engine.get_plan) is reused call after call.  The fingerprint holds
        # the identity of every array and a checksum of its contents (an in-place edit is noticed).
        fp = self._fingerprint()
        memo = getattr(self, "_block_memo", None)
        if memo is not None and memo[0] == fp:
            return memo[1], memo[2]
        E = len(self)
        lens = [len(f.xs) for f in self.dataframes]
        max_ind = int(np.max(lens))
        full = min(lens) == max_ind       # rows of one length (the usual case): nothing to pad, no fill pass
        make = np.empty if full else np.zeros
        xs = make((E, max_ind))
        ys = make((E, max_ind))
        yivar = make((E, max_ind))
        mask = np.empty((E, max_ind), dtype=bool) if full else np.ones((E, max_ind), dtype=bool)
        for i, f in enumerate(self.dataframes):
            n = lens[i]
            xs[i, :n] = f.xs
            ys[i, :n] = f.ys
            yivar[i, :n] = f.yivar
            mask[i, :n] = f.mask
        index_span = np.arange(0, E, dtype=int)
        metablock = {"index": index_span}
        for key in self.metadata:
            if key in self.metakeys:
                epoch_indices = np.zeros(index_span.shape)
                for i, ele in enumerate(self.metakeys[key]):
                    epoch_indices[np.asarray(self.metadata[key]) == ele] = i
                epoch_uniques = self.metakeys[key]
            else:
                epoch_uniques, epoch_indices = np.unique(self.metadata[key], return_inverse=True)
            self.metakeys[key] = epoch_uniques
            metablock[key] = np.asarray(epoch_indices).astype(int)
        out = DataBlock(xs=xs, ys=ys, yivar=yivar, mask=mask), DataBlock(**metablock)
        self._block_memo = (fp, out[0], out[1])
        return out

    def _fingerprint(self):
        """Identity of every frame array + a checksum of its bytes.  The pointer table of the arrays is
        kept while the array OBJECTS stay the same (then only the checksums are recomputed: one library
        call, OpenMP over the arrays -- jb_host_checksum, include/jabble_b200.h)."""
        import ctypes
        from . import _lib
        arrays = [a for f in self.dataframes for a in (f.xs, f.ys, f.yivar, f.mask)]
        cache = getattr(self, "_fp_cache", None)
        if cache is None or len(cache[0]) != len(arrays) or any(a is not b for a, b in zip(arrays, cache[0])):
            keep, ptrs, nbytes, shapes = [], [], [], []
            for a in arrays:
                arr = a if isinstance(a, np.ndarray) and a.flags.c_contiguous else np.ascontiguousarray(a)
                keep.append(arr)
                ptrs.append(arr.ctypes.data if arr.size else 0)
                nbytes.append(arr.nbytes)
                shapes.append((id(a), arr.shape, arr.dtype.str))
            direct = all(k is a for k, a in zip(keep, arrays))       # every frame array is a contiguous ndarray
            cache = (arrays, np.asarray(ptrs, dtype=np.uint64), np.asarray(nbytes, dtype=np.int64), tuple(shapes), direct)
            self._fp_cache = cache if direct else None               # copies were made: their pointers describe this call only
        _, ptrs, nbytes, shapes, _ = cache
        sums = np.empty(len(arrays), dtype=np.uint64)
        if len(arrays):
            _lib.check(_lib.load().jb_host_checksum(ctypes.c_void_p(ptrs.ctypes.data), ctypes.c_void_p(nbytes.ctypes.data),
                                                    len(arrays), ctypes.c_void_p(sums.ctypes.data)))
        parts = [shapes, sums.tobytes()]
        for key in sorted(self.metadata):
            arr = np.asarray(self.metadata[key])
            parts.append((key, id(self.metadata[key]), arr.shape, arr.tobytes() if arr.size <= 4096 else hash(arr.tobytes())))
        return tuple(parts)
