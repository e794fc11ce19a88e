"""Import shim: the spec's package directory name contains hyphens, so expose it as `flic_b200`."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module("fast-losless-image-compression-format_b200")
globals().update({k: v for k, v in vars(_pkg).items() if not k.startswith("__")})
