"""Import shim: the package directory is named `fermi.jl_b200/` (not a valid Python identifier), so
`import fermi_jl_b200` loads it from there and exposes its submodules (fermi_jl_b200.synth, ...)."""
import importlib.util as _ilu
import os as _os
import sys as _sys

_dir = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "fermi.jl_b200")
_spec = _ilu.spec_from_file_location(__name__, _os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = _ilu.module_from_spec(_spec)
_sys.modules[__name__] = _mod
_spec.loader.exec_module(_mod)
