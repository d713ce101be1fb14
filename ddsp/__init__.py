"""Drop-in alias: ``import ddsp`` resolves to the B200 implementation with the reference's module layout
(ddsp/__init__.py:1-4 of hugofloresgarcia/ddsp_pytorch: ``from .core import *`` + ``models``), so call
sites such as ``ddsp.harmonic_synth``, ``ddsp.models.decoder.DDSPDecoder`` or
``from ddsp.core import multiscale_fft`` keep working unchanged.  ``ddsp.utils`` / ``ddsp.data``
(matplotlib / lightning helpers) are outside the hot path and are not provided."""
import sys as _sys

import ddsp_pytorch_b200 as _impl
from ddsp_pytorch_b200.core import *  # noqa: F401,F403
from ddsp_pytorch_b200 import core, models  # noqa: F401

for _name in ("core", "models", "models.modules", "models.decoder", "models.encoder"):
    _sys.modules[__name__ + "." + _name] = _sys.modules["ddsp_pytorch_b200." + _name]
