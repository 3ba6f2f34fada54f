"""Importable alias of the package directory ``quantum-walks-time-dependent-hamiltonians_b200``
(hyphens are not valid in an ``import`` statement)."""

import importlib
import os
import sys

_ROOT = os.path.dirname(os.path.abspath(__file__))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)
_pkg = importlib.import_module("quantum-walks-time-dependent-hamiltonians_b200")
sys.modules[__name__] = _pkg
