"""Import shim: the package directory is named `networkdynamics.jl_b200` (not a valid Python identifier), so
`import ndb200` loads it under the module name `networkdynamics_jl_b200` and aliases it."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "networkdynamics.jl_b200")
_name = "networkdynamics_jl_b200"
if _name not in sys.modules:
    _spec = importlib.util.spec_from_file_location(_name, os.path.join(_dir, "__init__.py"),
                                                   submodule_search_locations=[_dir])
    _mod = importlib.util.module_from_spec(_spec)
    sys.modules[_name] = _mod
    _spec.loader.exec_module(_mod)
sys.modules[__name__] = sys.modules[_name]
