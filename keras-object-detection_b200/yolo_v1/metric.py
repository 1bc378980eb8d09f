"""Module-name shim: lets the reference's scripts keep `from metric import ...` (they run with
yolo_v1/ as the working directory) while the implementation is yolohot.metric on libyolohot."""
import os as _os
import sys as _sys

_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
from yolohot.metric import *  # noqa: F401,F403,E402
from yolohot.metric import __all__  # noqa: F401,E402
