"""`import avzoom` -> the package in ./real-time-audio-visual-zooming_b200 (a directory name Python's
import statement cannot spell).  Submodules resolve through the real package's __path__."""
import importlib as _importlib
import os as _os
import sys as _sys

_root = _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))
if _root not in _sys.path:
    _sys.path.insert(0, _root)
_pkg = _importlib.import_module("real-time-audio-visual-zooming_b200")
_sys.modules[__name__] = _pkg
