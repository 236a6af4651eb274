"""Empty stub: imported by the reference, never used on the hot path."""
