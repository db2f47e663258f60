"""Name shim: the reference package surface (src/mobrob/__init__.py:1-4) on the CUDA path."""
from mobrob_b200 import get_env, load_policy  # noqa: F401

__all__ = ["get_env", "load_policy"]
