"""Import shim for the upstream reference (earthspecies/avex) -- OFFLINE TOOLING ONLY.

Used by tests/golden/make_golden.py (and nothing else) to import the Python reference from
/root/reference inside the build container so that golden vectors can be generated.  It is never
imported by the product package, by `-m gpu` tests, by bench.py or by smoke(): /root/reference does
not exist on the GPU box.

The reference needs three modules that are absent here (gcsfs, s3fs, h5py: cloud IO / HDF5, unused
on the hot path) and an installed dist-info for `importlib.metadata.version("avex")`.
"""
from __future__ import annotations

import importlib.metadata as _md
import sys
import types

REF_ROOT = "/root/reference"


def install(ref_root: str = REF_ROOT):
    for name in ("gcsfs", "s3fs", "h5py"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.GCSFileSystem = type("GCSFileSystem", (), {})
            m.S3FileSystem = type("S3FileSystem", (), {})
            sys.modules[name] = m
    _orig = _md.version

    def _version(n):
        if n == "avex":
            return "0.0.0+ref"
        return _orig(n)

    _md.version = _version
    if ref_root not in sys.path:
        sys.path.insert(0, ref_root)
    import avex  # noqa: F401

    return avex
