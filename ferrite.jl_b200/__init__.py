"""ferrite_b200: B200-native global FE assembly behind Ferrite.jl's API surface.

The directory is named after the reference (`ferrite.jl_b200/`); import it as `ferrite_b200`
(the repo-root shim `ferrite_b200.py` registers it under that name).
"""
from ._lib import FB2Error, DetJNotPositive, MissingPatternEntry, LIB_PATH, declared_symbols, lib  # noqa: F401
from .api import *  # noqa: F401,F403
from . import api as _api

__all__ = [n for n in dir(_api) if not n.startswith("_")]
