"""Unpickle stub for `chumpy` (absent here). TEST INFRASTRUCTURE ONLY.

The FLAME pickle (/root/reference/head_detector/generic_model.pkl) stores `shapedirs`
as a pickled `chumpy.ch.Ch`; only its ndarray payload (state key `x`) is needed.
"""
from . import ch  # noqa: F401
