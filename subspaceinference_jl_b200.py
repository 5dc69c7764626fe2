"""Import alias: the package directory is named `subspaceinference.jl_b200/` (not a valid
Python identifier), so this module loads it and re-exports it as `subspaceinference_jl_b200`."""
import importlib.util
import sys
from pathlib import Path

_pkg_dir = Path(__file__).resolve().parent / "subspaceinference.jl_b200"
_spec = importlib.util.spec_from_file_location(__name__, _pkg_dir / "__init__.py", submodule_search_locations=[str(_pkg_dir)])
_mod = importlib.util.module_from_spec(_spec)
sys.modules[__name__] = _mod
_spec.loader.exec_module(_mod)
