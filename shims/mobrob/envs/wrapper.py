"""Name shim: ``from mobrob.envs.wrapper import get_env`` (src/mobrob/__init__.py:1)."""
from mobrob_b200.envs.wrapper import *  # noqa: F401,F403
from mobrob_b200.envs.wrapper import get_env  # noqa: F401
