"""Importable alias of the package directory (its mandated name contains hyphens).

``import segnn_b200`` is the package itself, and ``import segnn_b200.generic`` (any submodule) is the SAME module object
as the submodule under the package's real name: without that, the import system would execute a second copy of the
file under the alias name, and state set through one name (module globals, caches) would be invisible to the other."""
import importlib
import importlib.abc
import importlib.util
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
PACKAGE_NAME = "extending-the-n-body-benchmark-a-cross-model-study-of-geometric-deep-learning-architectures_b200"
_ALIAS = __name__


class _AliasFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    """Resolves ``segnn_b200.<sub>`` to the module ``<real package>.<sub>`` (imported on demand)."""

    def find_spec(self, fullname, path=None, target=None):
        if fullname.startswith(_ALIAS + "."):
            return importlib.util.spec_from_loader(fullname, self)
        return None

    def create_module(self, spec):
        return importlib.import_module(PACKAGE_NAME + spec.name[len(_ALIAS):])

    def exec_module(self, module):  # already executed under its real name
        return None


if not any(isinstance(f, _AliasFinder) for f in sys.meta_path):
    sys.meta_path.insert(0, _AliasFinder())
_pkg = importlib.import_module(PACKAGE_NAME)
sys.modules[__name__] = _pkg
