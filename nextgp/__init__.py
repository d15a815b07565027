"""Import shim: the package directory is literally named `nextgp.jl_b200/` (a dot is not importable as
one module name), so `import nextgp.jl_b200` resolves through this namespace package."""
import importlib.util as _u
import os as _os
import sys as _sys

_dir = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "nextgp.jl_b200")
if "nextgp.jl_b200" not in _sys.modules:
    _spec = _u.spec_from_file_location("nextgp.jl_b200", _os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
    jl_b200 = _u.module_from_spec(_spec)
    _sys.modules["nextgp.jl_b200"] = jl_b200
    _spec.loader.exec_module(jl_b200)
else:
    jl_b200 = _sys.modules["nextgp.jl_b200"]
