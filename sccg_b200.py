"""Import shim: the package directory is named `sccg-genome-compression_b200` (not a valid Python
identifier), so it is loaded here under the module name `sccg_b200`."""
import importlib.util
import sys
from pathlib import Path

_pkg = Path(__file__).resolve().parent / "sccg-genome-compression_b200"
_spec = importlib.util.spec_from_file_location("sccg_genome_compression_b200", _pkg / "__init__.py", submodule_search_locations=[str(_pkg)])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["sccg_genome_compression_b200"] = _mod
_spec.loader.exec_module(_mod)
globals().update({k: v for k, v in vars(_mod).items() if not k.startswith("__")})
