from . import particle  # noqa: F401
from .result import FilterResult  # noqa: F401
