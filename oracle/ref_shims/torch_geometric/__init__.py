"""Stand-in for ``torch_geometric`` (2.6.1 API subset) -- TEST INFRASTRUCTURE, see ../README.md."""
__version__ = "2.6.1+segnn_b200_shim"
IS_SHIM = True
from . import data, nn, loader  # noqa: E402,F401
