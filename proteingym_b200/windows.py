"""Window arithmetic for sequences longer than the model context (host side).
Reference: proteingym/utils/scoring_utils.py:43-52 (get_optimal_window), used at compute_fitness.py:492-495."""
from __future__ import annotations

import numpy as np


def optimal_window_starts(positions: np.ndarray, n_tokens: int, model_window: int = 1024):
    """Vectorised ``get_optimal_window``: for every masked token index return the start of its window and the common
    window length. Windows are ``[start, start + T)`` with ``T = min(n_tokens, model_window)``:
    left-aligned while the position is in the first half-window, right-aligned in the last half-window, centred
    (``pos - model_window // 2``) otherwise."""
    positions = np.asarray(positions, dtype=np.int64)
    if n_tokens <= model_window:
        return np.zeros(len(positions), dtype=np.int32), int(n_tokens)
    half = model_window // 2
    starts = positions - half
    starts = np.where(positions < half, 0, starts)
    starts = np.where(positions >= n_tokens - half, n_tokens - model_window, starts)
    return starts.astype(np.int32), int(model_window)
