#!/usr/bin/env python3
"""README name of the correction step (reference README.md:112-113,137-138): same program as badger.py, same top level."""
from badger import cli

if __name__ == "__main__":
    cli()
