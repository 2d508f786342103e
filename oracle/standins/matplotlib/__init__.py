"""Stub so that ``pyfilter.inference.plot`` imports; plotting is out of scope."""
from . import pyplot, axes  # noqa: F401
