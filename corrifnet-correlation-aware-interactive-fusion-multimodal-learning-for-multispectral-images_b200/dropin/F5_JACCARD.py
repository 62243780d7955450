"""Drop-in for the reference's F5_JACCARD.py (Jaccard only)."""
import os as _os
import sys as _sys

_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))))
from corrif_b200.metrics import Jaccard  # noqa: E402,F401
