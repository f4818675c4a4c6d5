"""Import shim: the product package lives in the directory `rust-msbwt_b200/`, whose name
is not a Python identifier.  `import rust_msbwt_b200` loads that directory as a package
under this module's name."""
import importlib.util
import os
import sys

_pkg_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "rust-msbwt_b200")
_spec = importlib.util.spec_from_file_location(
    __name__, os.path.join(_pkg_dir, "__init__.py"), submodule_search_locations=[_pkg_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules[__name__] = _mod
_spec.loader.exec_module(_mod)
