"""KmerIndexer with the reference's interface (algbio/Badger barcode_extraction/kmer_indexer.py:10-75).

``get_occurrences`` counts, for a query, the shared k-mer position pairs with every known string and
returns the best ones in the reference's dict order.  The counting (kmer_indexer.py:52-55) is the Q x W
scoring kernel ``bdg_kmer_score``; the selection rules (kmer_indexer.py:57-75: minimum count, hits_delta below
the best, stable order by count, max_hits, one entry per string) are array operations on the host.  The GPU path covers what the barcode hot path needs - 16-bp strings, k = 6
(SURVEY.md §8 a-5); other shapes (the R1-adapter search of the extraction step, which is out of scope)
raise NotImplementedError rather than fall back to a CPU implementation.
"""
from __future__ import annotations

import numpy as np

from . import ops


def _is_bc(s) -> bool:
    return len(s) == 16 and not (set(s) - set("ACGT"))


class KmerIndexer:
    def __init__(self, known_strings, kmer_size=6):
        self.seq_list = list(known_strings)
        self.k = kmer_size
        self._packed = None
        self._words = None

    def _get_kmers(self, seq):
        """The k-mers of seq from left to right (kmer_indexer.py:20-27)."""
        for i in range(len(seq) - self.k + 1):
            yield seq[i:i + self.k]

    def append(self, barcode):
        self.seq_list.append(barcode)
        self._packed = None

    def empty(self):
        return len(self.seq_list) == 0

    def _require_gpu_shape(self, sequence):
        if self.k != 6 or not _is_bc(sequence) or not all(_is_bc(s) for s in self.seq_list):
            raise NotImplementedError("KmerIndexer on the B200 path scores 16-bp ACGT strings with k=6; "
                                      "other shapes belong to the extraction step, which is out of scope")

    def get_occurrences(self, sequence, max_hits=0, min_kmers=1, hits_delta=1, ignore_equal=False):
        """kmer_indexer.py:49-75: {string: (string, shared k-mer count, query positions)}."""
        if not self.seq_list:
            return {}
        self._require_gpu_shape(sequence)
        if self._packed is None:                      # (re)build the device-resident index after a change of the string list
            self._words = ops.pack16(self.seq_list)[0]
            self._packed = ops.KmerIndex(self._words)
        q = ops.pack16([sequence])[0]
        _, hw, cnt, mult = self._packed.query(q, min_kmers=1)
        if hw.size == 0:
            return {}
        # the reference's dict order is first touch: by the first query position with a match, then by place in the string list
        order = np.lexsort((hw, np.argmax(mult > 0, axis=1)))
        hw, cnt, mult = hw[order], cnt[order].astype(np.int64), mult[order]
        keep = cnt >= min_kmers
        if ignore_equal:
            keep &= self._words[hw] != q[0]            # valid 16-mers: equal words are equal strings
        if not keep.any():
            return {}
        keep &= cnt >= cnt[keep].max() - hits_delta
        sel = np.flatnonzero(keep)
        sel = sel[np.argsort(-cnt[sel], kind="stable")]
        if max_hits:
            sel = sel[:max_hits]
        out = {}
        for o in sel.tolist():                         # a string listed twice keeps its first place and its last value, as in the reference
            s = self.seq_list[int(hw[o])]
            out[s] = (s, int(cnt[o]), np.repeat(np.arange(11), mult[o]).tolist())
        return out


class ArrayKmerIndexer(KmerIndexer):
    """kmer_indexer.py:78-154: same results as KmerIndexer (unused in the reference)."""
