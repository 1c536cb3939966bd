"""CPU: host-side logic of avex_b200 that needs no GPU (tables, C-ABI surface, loud failure without CUDA)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

import avex_b200
from avex_b200 import _lib
from avex_b200.fbank import KaldiFbank

G = os.path.join(os.path.dirname(__file__), "golden")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_tables_bit_exact_with_reference():
    t = np.load(os.path.join(G, "fbank_tables.npz"))
    fb = KaldiFbank()
    assert np.array_equal(fb.window.numpy(), t["window"])
    assert np.array_equal(fb.mel_fb.numpy(), t["mel_fb"])
    assert set(fb.state_dict().keys()) == {"window", "mel_fb"}  # reference buffer names (beats.py:76,80)


def test_fbank_geometry_guard():
    with pytest.raises(ValueError):
        KaldiFbank(num_mel_bins=64)
    fb = KaldiFbank()
    for n in (4000, 8000, 16000, 32000, 160000):
        assert fb.num_frames(n) == 1 + (n - 400) // 160


def test_library_exports_every_declared_symbol():
    """include/avexk.h is the contract: every `avexk_*` function it declares must be exported and bound."""
    header = open(os.path.join(ROOT, "include", "avexk.h")).read()
    declared = set(re.findall(r"\b(avexk_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    lib = ctypes.CDLL(_lib.LIB_PATH)  # built by __graft_entry__.build() / python -m avex_b200.build
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in avexk.h but not exported by libavexk.so"
    _lib.load()
    assert _lib.load().avexk_version() >= 100
    assert _lib.load().avexk_fbank_num_frames(160000) == 998
    assert _lib.load().avexk_beats_num_tokens(160000) == 496


def test_no_cpu_fallback():
    fb = KaldiFbank()
    with pytest.raises(_lib.AvexkError):
        fb(torch.zeros(1, 16000))
