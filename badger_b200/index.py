"""QGramIndex with the reference's interface (algbio/Badger index.py:12-93), scored on the GPU.

The reference keeps 4 096 dicts {number: multiplicity}; ``get_close`` walks the query's eleven 6-mer
buckets and returns the numbers ``> number`` whose accumulated multiplicity reaches the threshold
(index.py:77-93).  That accumulated value is S(a,b) = #{(p,q): 6mer_a[p] == 6mer_b[q]}, which
``bdg_kmer_score`` evaluates for the query against every indexed barcode in one launch.
"""
from __future__ import annotations

import numpy as np

from . import ops
from .common import rank


class QGramIndex:

    RANK = {'A': 0, 'C': 1, 'G': 2, 'T': 3}

    def __init__(self, threshold, bc_len, q=2):
        self.q = q
        print("k:", self.q)                                   # index.py:21
        self.threshold = bc_len - q + 1 - q * threshold        # index.py:22-24
        if self.threshold <= 0:
            self.threshold = 4
        self.bc_len = bc_len
        self._packed = []       # packed barcode per entry
        self._numbers = []      # caller's number per entry (the rank, in the reference's use)
        self._arr = None

    def _adopt(self, ranks: np.ndarray):
        """Bulk form of add_to_index(unrank(r), r) for every r (barcode_graph.py:204)."""
        self._packed = np.asarray(ranks, dtype=np.uint32).tolist()
        self._numbers = list(self._packed)
        self._arr = None

    def add_to_index(self, barcode, number):
        """index.py:29-35."""
        if self.q != 6 or len(barcode) != 16:
            raise NotImplementedError("the B200 path indexes 16-bp barcodes by 6-mers only")
        self._packed.append(rank(barcode, 16))
        self._numbers.append(number)
        self._arr = None

    def rank(self, qgram):
        """index.py:68-72."""
        r = 0
        for i in range(0, self.q):
            r += QGramIndex.RANK[qgram[i]] * (pow(4, i))
        return r

    def update_rank(self, rank_, b):
        """index.py:74-75."""
        return rank_ // 4 + QGramIndex.RANK[b] * (pow(4, self.q - 1))

    def get_close(self, barcode, number):
        """index.py:77-93: numbers j > number with S(barcode, entry_j) >= threshold (order unspecified)."""
        return self.get_close_many([barcode], [number])[0]

    def get_close_many(self, barcodes, numbers):
        """get_close for a batch of queries - the loop of barcode_graph.py:233-236 as ONE kernel launch over the resident index.
        Returns one list per query."""
        if len(barcodes) != len(numbers):
            raise ValueError("one number per barcode expected")
        if self.q != 6 or any(len(b) != 16 for b in barcodes):
            raise NotImplementedError("the B200 path indexes 16-bp barcodes by 6-mers only")
        if not self._packed or not len(barcodes):
            return [[] for _ in barcodes]
        if self._arr is None:                          # (re)build the device-resident index after add_to_index
            self._arr = ops.KmerIndex(np.asarray(self._packed, dtype=np.uint32))
        q = np.asarray([rank(b, 16) for b in barcodes], dtype=np.uint32)
        hq, hw, _, _ = self._arr.query(q, min_kmers=self.threshold)
        o = np.lexsort((hw, hq))
        hq, hw = hq[o], hw[o]
        first = np.searchsorted(hq, np.arange(len(barcodes) + 1))
        res = []
        for i, number in enumerate(numbers):
            out = {}
            for w in hw[first[i]:first[i + 1]].tolist():
                j = self._numbers[w]
                if j > number:
                    out[j] = True      # duplicates of a number collapse like dict keys do in the reference
            res.append(list(out))
        return res
