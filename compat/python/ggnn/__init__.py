"""`import ggnn` for programs written against the reference's Python module (python-src/ggnn/__init__.py,
src/ggnn/python/nanobind.cu:131-301): put <repo>/compat/python (and <repo>) on PYTHONPATH and the same script runs on
the B200 implementation.  Everything lives in ggnn_b200; this package only re-exports it under the reference's name."""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from ggnn_b200 import *  # noqa: E402,F401,F403
from ggnn_b200 import __all__ as _names  # noqa: E402

__all__ = list(_names)
__version__ = "0.9.0+b200"
