"""The files either side of the hot path, read and written by the native library:

    ExtractionTsv      badger.py:91-111   pandas.read_csv(sep="\\t") + the per-row loop that builds `read_assignment`
    .write             barcode_graph.py:388-410   `<out>_output_file.tsv` through DataFrame.to_csv(sep="\\t", index=False)
    whitelist_records  badger.py:82-88    `set(open(path).read().split("\\n"))`

The reader reproduces the subset of pandas' behaviour that extraction TSVs exercise and refuses any other file
(`Unsupported`): the caller then reads it through pandas, as the reference does.  Host code; no GPU involved.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import check, lib, ptr

ROW_BARCODE, ROW_EMIT = 1, 2
NONE = np.uint64(1) << np.uint64(32)          # == pipeline.NONE: "no centre", written as '*'


class Unsupported(Exception):
    """The native reader refuses the file (quotes, carriage returns, ids pandas would re-type, ...)."""


class ExtractionTsv:
    def __init__(self, path: str, bc_len: int = 16, threads: int = 0):
        h = C.c_void_p()
        rc = lib().bdg_tsv_open(str(path).encode(), int(bc_len), int(threads), C.byref(h))
        if rc == _lib.BDG_ERR_UNSUPPORTED:
            raise Unsupported((lib().bdg_last_error() or b"").decode(errors="replace"))
        check(rc)
        self._h = h
        self.rows = int(lib().bdg_tsv_rows(h))
        self.seqs16 = np.empty((self.rows, 16), np.uint8)
        self.kind = np.empty(self.rows, np.uint8)
        check(lib().bdg_tsv_barcodes(h, ptr(self.seqs16), ptr(self.kind)))

    @property
    def has_barcode(self) -> np.ndarray:
        """Rows whose barcode goes into the graph (badger.py:97-101 + barcode_graph.py:195-197)."""
        return (self.kind & ROW_BARCODE) != 0

    @property
    def emitted(self) -> np.ndarray:
        """Rows that get a line in the output file (badger.py:103-110)."""
        return (self.kind & ROW_EMIT) != 0

    def write(self, out_path: str, centre_per_row: np.ndarray, threads: int = 0) -> None:
        c = np.ascontiguousarray(centre_per_row, dtype=np.uint64)
        if c.size != self.rows:
            raise ValueError("one centre per TSV row expected (%d rows, got %d)" % (self.rows, c.size))
        check(lib().bdg_tsv_write_assignments(self._h, str(out_path).encode(), ptr(c), int(threads)))

    def write32(self, out_path: str, centre_per_row: np.ndarray, has_centre: np.ndarray, threads: int = 0) -> None:
        """write() from the 5-byte form of the per-row result (ops.assign_reads32)."""
        c = np.ascontiguousarray(centre_per_row, dtype=np.uint32)
        h = np.ascontiguousarray(has_centre, dtype=np.uint8)
        if c.size != self.rows or h.size != self.rows:
            raise ValueError("one centre per TSV row expected (%d rows, got %d / %d)" % (self.rows, c.size, h.size))
        check(lib().bdg_tsv_write_assignments32(self._h, str(out_path).encode(), ptr(c), ptr(h), int(threads)))

    def close(self):
        if self._h is not None:
            lib().bdg_tsv_close(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def whitelist_records(path: str) -> np.ndarray:
    """uint8[W, 16]: the entries of the whitelist file that are exactly 16 characters long (letters unchecked)."""
    h = C.c_void_p()
    check(lib().bdg_lines16_open(str(path).encode(), C.byref(h)))
    try:
        n = int(lib().bdg_lines16_count(h))
        out = np.empty((n, 16), np.uint8)
        if n:
            C.memmove(out.ctypes.data, lib().bdg_lines16_data(h), n * 16)
    finally:
        lib().bdg_lines16_close(h)
    return out
