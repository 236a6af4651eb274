"""Empty stub for matplotlib.pyplot (reference barcode_graph.py:12)."""
