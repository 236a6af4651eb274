"""KmerIndexer with the reference's interface (algbio/Badger barcode_extraction/kmer_indexer.py:10-75).

``get_occurrences`` counts, for a query, the shared k-mer position pairs with every known string and
returns the best ones in the reference's dict order.  The counting (kmer_indexer.py:52-55) is the Q x W
scoring kernel ``bdg_kmer_score``; the selection logic (kmer_indexer.py:57-75) is host code that follows the
reference line by line.  The GPU path covers what the barcode hot path needs - 16-bp strings, k = 6
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

    def _get_kmers(self, seq):
        """kmer_indexer.py:20-27."""
        if len(seq) < self.k:
            return
        kmer = seq[:self.k]
        yield kmer
        for i in range(self.k, len(seq)):
            kmer = kmer[1:] + seq[i]
            yield kmer

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
            self._packed = ops.KmerIndex(ops.pack16(self.seq_list)[0])
        q = ops.pack16([sequence])[0]
        _, hw, cnt, mult = self._packed.query(q, min_kmers=1)
        if hw.size == 0:
            return {}
        first_pos = np.argmax(mult > 0, axis=1)
        order = np.lexsort((hw, first_pos))            # dict insertion order of the reference: first touch
        result = []
        for o in order.tolist():
            i, count = int(hw[o]), int(cnt[o])
            if count < min_kmers:
                continue
            if ignore_equal and self.seq_list[i] == sequence:
                continue
            positions = [p for p in range(11) for _ in range(int(mult[o, p]))]
            result.append((self.seq_list[i], count, positions))
        if not result:
            return {}
        top_hits = max(result, key=lambda x: x[1])[1]
        result = filter(lambda x: x[1] >= top_hits - hits_delta, result)
        result = sorted(result, reverse=True, key=lambda x: x[1])
        if max_hits == 0:
            return {x[0]: x for x in result}
        return {x[0]: x for x in list(result)[:max_hits]}


class ArrayKmerIndexer(KmerIndexer):
    """kmer_indexer.py:78-154: same results as KmerIndexer (unused in the reference)."""
