"""Native op wrappers (see native.py)."""
from .native import *  # noqa: F401,F403
from . import native  # noqa: F401
