"""Import shim: makes the on-disk package directory ``classicalspinmc.jl_b200/`` importable as
``classicalspinmc.jl_b200`` (a directory name with a dot cannot be imported directly)."""
import importlib.util as _ilu
import os as _os
import sys as _sys

_pkg_dir = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                         "classicalspinmc.jl_b200")
_name = __name__ + ".jl_b200"
if _name not in _sys.modules:
    _spec = _ilu.spec_from_file_location(_name, _os.path.join(_pkg_dir, "__init__.py"),
                                         submodule_search_locations=[_pkg_dir])
    _mod = _ilu.module_from_spec(_spec)
    _sys.modules[_name] = _mod
    _spec.loader.exec_module(_mod)
jl_b200 = _sys.modules[_name]
