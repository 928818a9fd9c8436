"""Import shim: the package directory is `cuda-pt_b200/` (hyphenated by the project layout), which
Python cannot import by name.  `import cuda_pt_b200` loads that directory as a regular package."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cuda-pt_b200")
_spec = importlib.util.spec_from_file_location("cuda_pt_b200", os.path.join(_dir, "__init__.py"),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["cuda_pt_b200"] = _mod
_spec.loader.exec_module(_mod)
