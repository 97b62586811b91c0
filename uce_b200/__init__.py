"""Importable alias of the ``unified-concept-editing_b200`` package directory.

The directory name required by the build layout contains hyphens, which Python cannot
import; this shim points ``uce_b200``'s search path at it so ``uce_b200.solver`` etc.
resolve to ``unified-concept-editing_b200/solver.py``.
"""
import os as _os

_REAL = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "unified-concept-editing_b200")
__path__.insert(0, _REAL)

from ._version import __version__  # noqa: E402,F401
