"""ddsp_pytorch_b200: the DDSP synthesis hot path on hand-written sm_100a CUDA kernels, behind the
Python API of hugofloresgarcia/ddsp_pytorch (``ddsp.core`` functions, ``ddsp.models`` modules)."""
from ._lib import load as _load

_load()

from .core import *          # noqa: E402,F401,F403  (mirrors ddsp/__init__.py:1)
from . import core           # noqa: E402,F401
from . import functions      # noqa: E402,F401
from . import models         # noqa: E402,F401

__version__ = "0.1.0"
