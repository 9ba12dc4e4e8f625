#!/usr/bin/env python
"""`python shgyield.py <input.yml>`: command line of the SSHG yield path (see shgyield_b200/cli.py)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from shgyield_b200.cli import main  # noqa: E402

if __name__ == "__main__":
    sys.exit(main())
