"""Module-name shim for the label half of the reference's yolo_v1/dataset.py (see yolohot.dataset)."""
import os as _os
import sys as _sys

_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
from yolohot.dataset import *  # noqa: F401,F403,E402
from yolohot.dataset import __all__  # noqa: F401,E402
