"""Name shim for the parts of gymnasium 0.28.1 the reference scripts touch."""
from . import spaces, wrappers  # noqa: F401

__version__ = "0.28.1"
