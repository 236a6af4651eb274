"""Stub for `ssw-py` (un-vendored). Only imported, never executed, on the
hot path (reference common.py:9, barcode_extraction/common.py:7)."""


class AlignmentMgr:
    def __init__(self, *a, **k):
        pass

    def __getattr__(self, name):
        raise NotImplementedError("ssw stub: %s is outside the oracle's scope" % name)
