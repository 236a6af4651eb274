#!/usr/bin/env python3
"""README name of the correction step (reference README.md:112-113,137-138): same program as badger.py."""
import sys

from badger import main

if __name__ == "__main__":
    main(sys.argv[1:])
