"""Stand-in for the ``e3nn`` package (0.5.1 API subset) -- TEST INFRASTRUCTURE, see ../README.md."""
__version__ = "0.5.1+segnn_b200_shim"
IS_SHIM = True
from . import o3, nn  # noqa: E402,F401
