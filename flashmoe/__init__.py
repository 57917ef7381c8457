"""`import flashmoe` -- the reference's package name, served by flashmoe_b200.

`flashmoe._C` is the COMPILED pybind11 extension built from csrc/python_bindings.cu (the same six functions as the
reference's extension, over the C-ABI of include/flashmoe_b200.h); if it has not been built the ctypes module
flashmoe_b200._C (same surface) takes its place."""
import sys as _sys

import flashmoe_b200 as _impl
from flashmoe_b200 import get_compiled_config, run_moe  # noqa: F401
from flashmoe_b200 import launcher, ops, worker  # noqa: F401

try:
    from . import _C  # noqa: F401  (flashmoe/_C.cpython-*.so)
except ImportError:
    try:
        from flashmoe_b200 import _C  # noqa: F401

        _sys.modules[__name__ + "._C"] = _C
    except Exception:  # the warning was already issued by flashmoe_b200
        pass
for _m in ("ops", "launcher", "worker"):
    _sys.modules[__name__ + "." + _m] = getattr(_impl, _m)

__version__ = _impl.__version__
__all__ = ["run_moe", "get_compiled_config"]
