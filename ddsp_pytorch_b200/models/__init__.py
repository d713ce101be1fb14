"""Mirrors ddsp/models/__init__.py."""
from . import modules  # noqa: F401
from . import decoder  # noqa: F401
from . import encoder  # noqa: F401
