"""String <-> packed helpers with the reference's names (algbio/Badger common.py:11-38)."""
from __future__ import annotations

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
