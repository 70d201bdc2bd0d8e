"""glpk.js_b200 -- B200-native drop-in for the simplex hot path of glpk.js.

This package is the host side above the C ABI of ``libglpb200.so``
(``include/glpb200.h``).  It holds

* ``native``   -- the ctypes binding of every C-ABI entry point;
* ``glpk``     -- a mirror of the reference's JavaScript API for this path
  (``glp_create_prob``, ``glp_read_lp``, ``glp_simplex``, ``glp_intopt``,
  ``SMCP``, ``IOCP``, ``glp_get_obj_val``, ``glp_get_col_prim``, status codes),
  same names, argument meaning and error behaviour as ``lib/glpapi*.js``.

The directory name contains a dot, so it is imported through the small
``glpk_js_b200.py`` shim at the repository root (``import glpk_js_b200``).

There is no CPU fallback: importing works without a GPU (so that the build
and symbol checks can run anywhere) but every compute call raises unless the
CUDA library is present and a device is visible.
"""
from . import native  # noqa: F401
from .native import LIB_PATH, load, build  # noqa: F401
from . import glpk  # noqa: F401
from . import bnb  # noqa: F401

__all__ = ["native", "glpk", "load", "build", "LIB_PATH"]
