"""String <-> packed helpers with the reference's names (algbio/Badger common.py:11-38)."""
from __future__ import annotations

import numpy as np

RANK = {'A': 0, 'C': 1, 'G': 2, 'T': 3}
UNRANK = {0: 'A', 1: 'C', 2: 'G', 3: 'T'}


def rank(seq, length):
    """common.py:21-25 (single string; bulk packing goes through ops.pack16 on the GPU)."""
    r = 0
    for i in range(0, length):
        r += RANK[seq[i]] << (2 * i)      # KeyError on a non-ACGT base, as in the reference
    return r


def unrank(rk, length):
    """common.py:27-38."""
    return "".join(UNRANK[(rk >> (2 * i)) & 3] for i in range(length))


def sorted_unique(x: np.ndarray) -> np.ndarray:
    """Ascending distinct values (np.sort + neighbour mask: numpy 2.3's hash-based np.unique is ~100x slower on 10^6 keys)."""
    s = np.sort(np.asarray(x).ravel())
    if s.size < 2:
        return s
    keep = np.empty(s.size, bool)
    keep[0] = True
    np.not_equal(s[1:], s[:-1], out=keep[1:])
    return s[keep]
