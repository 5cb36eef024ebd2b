"""Importable alias of the product package.

The package directory is named after the reference
(``real-time-multi-object-detection---tracking-system_b200``), which is not a Python
identifier; ``import rtmodt_b200`` gives the same module objects (no second copy).
"""

import importlib
import sys

_REAL = "real-time-multi-object-detection---tracking-system_b200"
_pkg = importlib.import_module(_REAL)
for _name, _mod in list(sys.modules.items()):
    if _name == _REAL or _name.startswith(_REAL + "."):
        sys.modules[__name__ + _name[len(_REAL):]] = _mod
sys.modules[__name__] = _pkg
