"""Import alias: the package directory is named `isl-signlanguage-translation_b200`, which is not a legal
Python identifier, so `import isl_b200` loads it under this name."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "isl-signlanguage-translation_b200")
_spec = importlib.util.spec_from_file_location("isl_b200", os.path.join(_dir, "__init__.py"),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["isl_b200"] = _mod
_spec.loader.exec_module(_mod)
