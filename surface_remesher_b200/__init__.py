"""Import alias: the package directory is `surface-remesher_b200/` (not a valid Python
identifier), so `import surface_remesher_b200` loads it from there."""
import importlib.util as _u
import pathlib as _p
import sys as _s

_dir = _p.Path(__file__).resolve().parent.parent / "surface-remesher_b200"
_spec = _u.spec_from_file_location(__name__, _dir / "__init__.py", submodule_search_locations=[str(_dir)])
_mod = _u.module_from_spec(_spec)
_s.modules[__name__] = _mod
_spec.loader.exec_module(_mod)
