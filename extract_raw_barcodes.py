#!/usr/bin/env python3
"""Entry-point shim for the reference's extract_raw_barcodes.py (reference extract_raw_barcodes.py:360-380).

Barcode EXTRACTION (polyT search, R1-adapter SSW alignment, FASTQ/BAM readers) is outside the hot path this
repository replaces: it contains no barcode scoring and depends on ssw-py, pysam and
Biopython.  Run the reference's own script for that step and feed its TSV to badger.py / barcodes.py here.
"""
import sys

if __name__ == "__main__":
    sys.stderr.write(__doc__)
    sys.exit(2)
