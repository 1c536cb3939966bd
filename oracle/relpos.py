"""Oracle: T5-style bidirectional relative-position buckets of BEATs attention (numpy).

Test infrastructure -- see oracle/__init__.py.  Restates
avex/models/beats/backbone.py:438-473 (`_relative_positions_bucket`) and :475-492 (`compute_bias`).
bias[h, i, j] = table[bucket(j - i), h] is Toeplitz, so only the 2N-1 distinct offsets are kept.
"""
from __future__ import annotations

import math

import numpy as np


def relative_position_bucket(rel: np.ndarray, num_buckets: int = 320, max_distance: int = 800) -> np.ndarray:
    """backbone.py:438-473 with bidirectional=True.  `rel` = key_pos - query_pos (int64)."""
    rel = np.asarray(rel, dtype=np.int64)
    nb = num_buckets // 2  # backbone.py:453
    out = (rel > 0).astype(np.int64) * nb  # backbone.py:454
    a = np.abs(rel)
    max_exact = nb // 2  # backbone.py:459
    is_small = a < max_exact
    with np.errstate(divide="ignore"):
        # float32 log, then / python float, * python int -- all float32 tensor arithmetic in torch
        lg = np.log(a.astype(np.float32) / np.float32(max_exact))
        scaled = lg / np.float32(math.log(max_distance / max_exact)) * np.float32(nb - max_exact)
    scaled = np.where(is_small, np.float32(0), scaled)  # masked lanes (log 0 = -inf) never used
    large = max_exact + scaled.astype(np.int64)  # truncation toward zero, backbone.py:462-466
    large = np.minimum(large, nb - 1)  # backbone.py:467-470
    return out + np.where(is_small, a, large)  # backbone.py:472


def bias_vector(table: np.ndarray, n_tokens: int, num_buckets: int = 320, max_distance: int = 800) -> np.ndarray:
    """[H, 2N-1] float32 with vec[h, (j - i) + N - 1] = table[bucket(j - i), h]; backbone.py:475-492."""
    rel = np.arange(-(n_tokens - 1), n_tokens, dtype=np.int64)
    b = relative_position_bucket(rel, num_buckets, max_distance)
    return np.ascontiguousarray(np.asarray(table, dtype=np.float32)[b].T)
