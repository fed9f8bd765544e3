"""Importable alias for the package directory `highperformancecomputing-latticeboltzmannmethod_b200`
(its name is not a Python identifier).  `import lbm_b200` gives the same module."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module("highperformancecomputing-latticeboltzmannmethod_b200")
globals().update({k: v for k, v in vars(_pkg).items() if not k.startswith("__")})
sys.modules[__name__] = _pkg
