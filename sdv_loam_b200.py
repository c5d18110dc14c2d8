"""Import shim: the package directory is `sdv-loam_b200/` (hyphen), which Python cannot import by name.
`import sdv_loam_b200` loads that directory as the package `sdv_loam_b200`."""
import importlib.util as _u
import os as _os
import sys as _sys

_dir = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "sdv-loam_b200")
_spec = _u.spec_from_file_location("sdv_loam_b200", _os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = _u.module_from_spec(_spec)
_sys.modules["sdv_loam_b200"] = _mod
_spec.loader.exec_module(_mod)
