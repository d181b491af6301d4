"""Import alias: ``import cpb200`` == the package in ``channel-pruning_b200/`` (whose directory
name is not a valid Python identifier).  Sub-modules are aliased too, so
``from cpb200.lib.decompose import dictionary`` and ``cpb200.lib.cfgs`` share state with the
real package."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_REAL = "channel-pruning_b200"
_pkg = importlib.import_module(_REAL)
for _name, _mod in list(sys.modules.items()):
    if _name == _REAL or _name.startswith(_REAL + "."):
        sys.modules["cpb200" + _name[len(_REAL):]] = _mod
sys.modules[__name__] = _pkg
