"""Mutant-string parsing to flat int32 site arrays (host half of ``label_row``,
reference: proteingym/baselines/esm/compute_fitness.py:240-250)."""
from __future__ import annotations

import numpy as np

from .alphabet import ALPHABET


def parse_mutants(mutants, sequence: str, offset_idx: int = 1):
    """-> (site_row, site_wt, site_mt, row_offsets) int32 arrays in CSR form.

    ``site_row`` is the token index ``1 + idx`` (BOS shifts by one, compute_fitness.py:248-249). Raises
    ``AssertionError("The listed wildtype does not match the provided sequence")`` exactly where the reference does
    (:244), and ``ValueError`` / ``IndexError`` for malformed entries like the reference's ``int()`` / indexing would."""
    rows, wts, mts, offs = [], [], [], [0]
    get_idx = ALPHABET.get_idx
    for row in mutants:
        for mutation in str(row).split(":"):
            wt, idx, mt = mutation[0], int(mutation[1:-1]) - offset_idx, mutation[-1]
            assert sequence[idx] == wt, "The listed wildtype does not match the provided sequence"
            rows.append(1 + idx)
            wts.append(get_idx(wt))
            mts.append(get_idx(mt))
        offs.append(len(rows))
    as32 = lambda a: np.asarray(a, dtype=np.int32)
    return as32(rows), as32(wts), as32(mts), as32(offs)
