"""A dict-backed double of the small h5py surface the embedding cache uses (File, create_dataset, attrs, keys, __getitem__).

h5py is not installed in this image.  Putting this directory first on sys.path lets the tests drive (a) the HDF5 branch of
avex_b200.extract.save_embeddings_arrays and (b) the REFERENCE's own save_embeddings_arrays / load_embeddings_arrays
(avex/evaluation/embedding_utils.py:1433-1678) against the same container, so the two writers can be compared attr by attr
and each file can be read back by the other side's loader.  Like h5py, list attrs come back as numpy arrays.
Test infrastructure only.
"""
import pickle

import numpy as np


class _Attrs(dict):
    def __getitem__(self, k):
        v = dict.__getitem__(self, k)
        return np.asarray(v) if isinstance(v, (list, tuple)) else v

    def get(self, k, default=None):
        return self[k] if k in self else default


class File:
    def __init__(self, name, mode="r", **_kw):
        self.filename, self.mode = str(name), mode
        self._data, self.attrs = {}, _Attrs()
        if mode == "r":
            with open(self.filename, "rb") as fh:
                blob = pickle.load(fh)
            self._data, self.attrs = blob["data"], _Attrs(blob["attrs"])
            self.creation = blob["creation"]
        else:
            self.creation = {}

    def create_dataset(self, name, shape=None, dtype=None, data=None, **kw):
        arr = np.asarray(data) if data is not None else np.zeros(shape, dtype=dtype or np.float32)
        self._data[name] = arr
        self.creation[name] = dict(kw)
        return arr

    def keys(self):
        return self._data.keys()

    def __contains__(self, k):
        return k in self._data

    def __getitem__(self, k):
        return self._data[k]

    def close(self):
        if self.mode != "r":
            with open(self.filename, "wb") as fh:
                pickle.dump({"data": self._data, "attrs": dict(self.attrs), "creation": self.creation}, fh)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False
