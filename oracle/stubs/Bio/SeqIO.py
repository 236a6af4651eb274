"""Empty stub for Bio.SeqIO (reference extract_raw_barcodes.py:20)."""
