"""Import shim: the product package lives in the directory `gradientrobustmultiphysics.jl_b200/`
(the name the task fixes; it is not a valid Python identifier), so `import grmp_b200`
loads that directory as the package `grmp_b200`."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "gradientrobustmultiphysics.jl_b200")
_spec = importlib.util.spec_from_file_location("grmp_b200", os.path.join(_dir, "__init__.py"),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["grmp_b200"] = _mod
_spec.loader.exec_module(_mod)
