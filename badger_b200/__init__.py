"""badger_b200 - B200-native drop-in for the barcode edit-distance hot path of algbio/Badger.

``BarcodeGraph``, ``QGramIndex`` and ``KmerIndexer`` keep the reference's Python interface; the work runs in
hand-written CUDA kernels (sm_100a) behind the C ABI of ``csrc/libbadger_b200.so`` (include/badger_b200.h).
"""
from ._lib import BadgerB200Error, LIB_PATH, init, lib      # noqa: F401
from . import ops                                           # noqa: F401
from . import pipeline                                      # noqa: F401
from .barcode_graph import BarcodeGraph                      # noqa: F401
from .common import rank, unrank                             # noqa: F401
from .index import QGramIndex                                # noqa: F401
from .kmer_indexer import ArrayKmerIndexer, KmerIndexer      # noqa: F401

__all__ = ["BarcodeGraph", "QGramIndex", "KmerIndexer", "ArrayKmerIndexer", "rank", "unrank", "ops", "pipeline", "init", "lib",
           "BadgerB200Error", "LIB_PATH"]
