"""MF/load_data.py of the reference -> CSR-backed loaders (see pda_b200/data.py for the mirrored interface)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pda_b200.data import Data, Data2  # noqa: E402,F401
