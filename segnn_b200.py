"""Importable alias of the package directory (its mandated name contains hyphens)."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
PACKAGE_NAME = "extending-the-n-body-benchmark-a-cross-model-study-of-geometric-deep-learning-architectures_b200"
_pkg = importlib.import_module(PACKAGE_NAME)
sys.modules[__name__] = _pkg
