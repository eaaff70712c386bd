"""Import shim: `import b200ens` loads the package directory `differentialequations.jl_b200/`
(whose name is not a valid Python identifier) under the module name `b200ens`."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "differentialequations.jl_b200")
_spec = importlib.util.spec_from_file_location("b200ens", os.path.join(_dir, "__init__.py"),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["b200ens"] = _mod
_spec.loader.exec_module(_mod)
