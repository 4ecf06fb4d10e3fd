"""picasso_b200 -- B200-native (sm_100a) implementation of picasso's
single-molecule localization hot path behind picasso's own function signatures.

Modules mirror the reference package layout for the path only:
``gaussmle``, ``gausslq``, ``localize``, ``render``, ``imageprocess``,
``postprocess``.  All numerics run in hand-written CUDA (libpicasso_b200.so,
C ABI in include/picasso_b200.h); there is no CPU fallback.
"""
__version__ = "0.1.0"
