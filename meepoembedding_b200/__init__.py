"""meepo-b200: a B200-native (sm_100a) dynamic embedding table behind a thin C ABI.

Importing this package never falls back to a CPU implementation: `Table()`
loads meepoembedding_b200/libmeepo.so (CUDA) and raises if it is missing.
"""
from . import _capi as capi
from . import keygen
from ._capi import MeepoError, load_library, product_library
from .table import Table

__all__ = ["Table", "MeepoError", "capi", "load_library", "product_library"]
