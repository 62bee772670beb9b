"""smash_b200 -- B200-native forward / adjoint solver behind smash's own solver entry points.

Only the hot path lives here: ``smash_b200.solver`` mirrors the f90wrap modules the reference's Python layer
imports (``_mw_forward``, ``_mw_multiple_run``, ``_mwd_*``); the arithmetic runs in ``libsmash_b200.so``
(hand-written CUDA for sm_100a, C ABI in ``include/smash_b200.h``).
"""
from . import _lib  # noqa: F401
from .solver._mw_forward import forward, forward_b, hyper_forward, hyper_forward_b  # noqa: F401
from .solver._mw_multiple_run import compute_multiple_run  # noqa: F401

__version__ = "0.1.0"
