"""Import shim: loads the package directory ``co-detr-tensorrt_b200/`` (a name Python cannot import
directly) under the module name ``codetr_b200``."""
import importlib.util
import os
import sys

_pkg_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "co-detr-tensorrt_b200")
_spec = importlib.util.spec_from_file_location(
    "codetr_b200", os.path.join(_pkg_dir, "__init__.py"), submodule_search_locations=[_pkg_dir]
)
_module = importlib.util.module_from_spec(_spec)
sys.modules["codetr_b200"] = _module
_spec.loader.exec_module(_module)
