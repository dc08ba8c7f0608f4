"""Imports the UNMODIFIED reference (``/root/reference``) in this container -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Used by ``tests/golden/make_reference_golden.py`` (golden vectors), by ``tests/test_reference_live.py`` (only when
``/root/reference`` exists) and by ``bench.py --impl reference`` / ``cpu_baseline`` (which time the reference's own
SEGNN + rollout code on the host cores when it is importable).

The reference's SEGNN path imports e3nn / torch_geometric / torch_scatter at module top.  When those packages import,
they are used as they are (``kind = "reference"``).  Otherwise the stand-ins under ``oracle/ref_shims`` are put on
``sys.path`` (``kind = "reference+shims"``: the reference's own module code on restated third-party primitives, see
``oracle/ref_shims/README.md``).  Plot / logging packages the path never calls (matplotlib, plotly, torchmetrics,
jsonargparse, optuna) and the six other model families that ``models/__init__.py`` and ``utils/nbody_utils.py`` import
eagerly are replaced by inert stubs.
"""
from __future__ import annotations

import importlib
import importlib.machinery
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
SHIMS = os.path.join(HERE, "ref_shims")
DEFAULT_REF = os.environ.get("SEGNN_REFERENCE_ROOT", "/root/reference")


class _Stub(types.ModuleType):
    """A module whose every attribute is an inert callable/class; enough for ``import x.y as z`` and ``from x import A``."""

    def __init__(self, name):
        super().__init__(name)
        self.__path__ = []  # behaves as a package
        self.__spec__ = importlib.machinery.ModuleSpec(name, None, is_package=True)

    def __getattr__(self, item):
        if item.startswith("__"):
            raise AttributeError(item)
        full = f"{self.__name__}.{item}"
        if full in sys.modules:
            return sys.modules[full]
        return _Inert(item)


class _Inert:
    """Callable, subscriptable, iterable, usable as a base class or context manager; does nothing."""

    def __init__(self, name="inert", *a, **k):
        self._name = name

    def __call__(self, *a, **k):
        return _Inert(self._name)

    def __getattr__(self, n):
        if n.startswith("__"):
            raise AttributeError(n)
        return _Inert(n)

    def __getitem__(self, k):
        return _Inert(self._name)

    def __setitem__(self, k, v):
        pass

    def __iter__(self):
        return iter(())

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    def __mro_entries__(self, bases):
        return (object,)


class _Axes(_Inert):
    def hist(self, data, bins=10, **kwargs):
        import numpy as np
        counts, edges = np.histogram(np.asarray(data), bins=bins)
        return counts, edges, None


def _fake_pyplot() -> types.ModuleType:
    """``matplotlib.pyplot`` with just enough behaviour for the reference's macro functions, which compute their
    statistics inside plotting routines (datasets/nbody/visualization_utils.py): ``subplots`` returns real containers
    of inert axes, ``Axes.hist`` returns real counts."""
    import numpy as np
    mod = _Stub("matplotlib.pyplot")
    mod.rcParams = {}

    def subplots(nrows=1, ncols=1, **kwargs):
        if nrows * ncols == 1:
            return _Inert("fig"), _Axes("ax")
        axes = np.empty((nrows, ncols), dtype=object)
        for i in range(nrows):
            for j in range(ncols):
                axes[i, j] = _Axes("ax")
        return _Inert("fig"), (axes.reshape(-1) if 1 in (nrows, ncols) else axes)

    mod.subplots = subplots
    return mod


def _stub(name: str):
    parts = name.split(".")
    for k in range(1, len(parts) + 1):
        sub = ".".join(parts[:k])
        if sub not in sys.modules:
            sys.modules[sub] = _Stub(sub)
            if k > 1:
                setattr(sys.modules[".".join(parts[:k - 1])], parts[k - 1], sys.modules[sub])


def _importable(name: str) -> bool:
    try:
        importlib.import_module(name)
        return True
    except Exception:
        return False


def third_party_kind() -> str:
    """'reference' when the real e3nn / torch_geometric / torch_scatter import, else 'reference+shims'."""
    real = True
    for name in ("e3nn", "torch_geometric", "torch_scatter"):
        if not _importable(name) or getattr(sys.modules.get(name), "IS_SHIM", False):
            real = False
    return "reference" if real else "reference+shims"


_PACKAGE_DIRS = ("models", "models.segnn", "utils", "datasets", "datasets.nbody", "datasets.nbody.dataset",
                 "datasets.nbody_offline", "dataloaders", "helper_scripts", "training")


def setup(ref_root: str = DEFAULT_REF, force_shims: bool = False) -> str:
    """Prepare ``sys.path`` / ``sys.modules`` so that ``import models.segnn.segnn`` etc. resolve into ``ref_root``.
    Returns the kind ('reference' or 'reference+shims').  Raises FileNotFoundError when ``ref_root`` is absent."""
    if not os.path.isdir(os.path.join(ref_root, "models", "segnn")):
        raise FileNotFoundError(f"reference checkout not found at {ref_root}")
    repo_root = os.path.dirname(HERE)
    if repo_root not in sys.path:
        sys.path.insert(0, repo_root)  # the shims import oracle.segnn_oracle for parity-odd couplings
    need_shims = force_shims or not all(_importable(n) for n in ("e3nn", "torch_geometric", "torch_scatter"))
    if need_shims and SHIMS not in sys.path:
        for name in [m for m in sys.modules if m.split(".")[0] in ("e3nn", "torch_geometric", "torch_scatter")]:
            del sys.modules[name]
        sys.path.insert(0, SHIMS)
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.widgets", "matplotlib.lines", "matplotlib.animation",
                 "matplotlib.backends", "matplotlib.backends.backend_pdf", "mpl_toolkits", "mpl_toolkits.mplot3d",
                 "plotly", "plotly.graph_objects", "plotly.express", "plotly.subplots", "torchmetrics", "jsonargparse",
                 "optuna"):
        if not _importable(name):
            _stub(name)
    if isinstance(sys.modules.get("matplotlib.pyplot"), _Stub) \
            and "subplots" not in vars(sys.modules["matplotlib.pyplot"]):
        sys.modules["matplotlib.pyplot"] = _fake_pyplot()
        sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    # `models/__init__.py` and `utils/nbody_utils.py` import every model family; only SEGNN is on the path
    for pkg in _PACKAGE_DIRS:
        if pkg not in sys.modules:
            mod = types.ModuleType(pkg)
            mod.__path__ = [os.path.join(ref_root, *pkg.split("."))]
            mod.__spec__ = importlib.machinery.ModuleSpec(pkg, None, is_package=True)
            mod.__spec__.submodule_search_locations = mod.__path__
            sys.modules[pkg] = mod
            if "." in pkg:
                parent, child = pkg.rsplit(".", 1)
                setattr(sys.modules[parent], child, mod)
            if pkg == "models":  # `from models import SEGNN, PaiNN, ...` (utils/utils_train.py:5-11)
                def _models_getattr(name):
                    if name == "SEGNN":
                        return importlib.import_module("models.segnn.segnn").SEGNN
                    if name.startswith("__"):
                        raise AttributeError(name)
                    return _Inert(name)
                mod.__getattr__ = _models_getattr
    for name in ("models.equiformer_v2", "models.equiformer_v2.architecture",
                 "models.equiformer_v2.architecture.equiformer_v2_nbody", "models.ponita", "models.ponita.ponita_nbody",
                 "models.CGENN", "models.CGENN.nbody_cgenn", "models.graph_transformer",
                 "models.graph_transformer.graph_transformer_torch", "models.PaiNN", "models.PaiNN.PaiNN",
                 "models.egnn_mc", "models.egnn_mc.egnn_mc"):
        if name not in sys.modules:
            _stub(name)
    if ref_root not in sys.path:
        sys.path.append(ref_root)
    datagen = os.path.join(ref_root, "datasets", "nbody_offline", "datagen")
    if datagen not in sys.path:
        sys.path.append(datagen)  # system.py does `from physical_objects import ...`
    return third_party_kind()


def load_wigner(ref_root: str = DEFAULT_REF):
    """The e3nn ``wigner_D`` code + ``Jd.pt`` constants the reference vendors (models/equiformer_v2/architecture/
    wigner.py:1-43), loaded by file path so that the equiformer package itself is not imported."""
    path = os.path.join(ref_root, "models", "equiformer_v2", "architecture", "wigner.py")
    spec = importlib.util.spec_from_file_location("_ref_wigner", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def available(ref_root: str = DEFAULT_REF) -> bool:
    return os.path.isdir(os.path.join(ref_root, "models", "segnn"))
